#!/bin/bash
# Round-2 one-GPU pass D: the whole GPU suite on the current build (WarpImages operation patched, pyramid on slabs, ...),
# smoke, host-side phase trace of the drop-in, launch list of the bench command.
tag=${1:-r2d}; out=gpurun_out; mkdir -p $out
DCMA_PARITY_REPORT=$out/${tag}_parity.jsonl timeout 1500 python -m pytest tests -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log; tail -12 $out/${tag}_pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
./dicomautomaton_b200/host/_build/warp_operation_check_patched run 2>/dev/null | tail -2
DCMA_B200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --iters 50 --no-cpu-baseline --no-other-modes --dropin-steps 3 --e2e-steps 2 > $out/${tag}_trace_bench.json 2> $out/${tag}_trace_bench.err; echo "trace bench rc=$?"; tail -22 $out/${tag}_trace_bench.err
timeout 200 python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu-baseline --no-other-modes --no-dropin --e2e-steps 1 > /dev/null 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 1 --warmup 3 --iters 20 --no-cpu-baseline --no-other-modes --no-dropin --e2e-steps 1 > $out/${tag}_nculist.log 2>&1; echo "ncu list rc=$?"
