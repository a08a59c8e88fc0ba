#!/bin/bash
# One GPU-box pass: GPU test-suite, smoke, bench (both arms), ncu launch list.  Outputs under gpurun_out/.
# usage (from the repo root, under gpurun):  bash tools/gpu_check.sh [tag]
tag=${1:-run}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
nproc > $out/${tag}_nproc.txt
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
python bench.py --impl reference --steps 1 --warmup 0 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err
python bench.py > $out/${tag}_bench_fp64.json 2> $out/${tag}_bench_fp64.err
python bench.py --field-mode fp32 --no-cpu-baseline --no-other-modes > $out/${tag}_bench_fp32.json 2> $out/${tag}_bench_fp32.err
for mode in fp32 fp64; do
  python tools/profile_run.py --mode $mode --iters 3 > $out/${tag}_plain_$mode.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_$mode.csv \
      python tools/profile_run.py --mode $mode --iters 3 > $out/${tag}_ncu_$mode.log 2>&1
done
tail -3 $out/${tag}_pytest.log; tail -2 $out/${tag}_smoke.log; cat $out/${tag}_bench_fp64.json | cut -c1-400
