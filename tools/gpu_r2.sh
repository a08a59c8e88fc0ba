#!/bin/bash
# Round-2 one-GPU pass: parity tests on the default build, then the bench of each tuning variant.
# usage: bash tools/gpu_r2.sh <tag> "<variant tags>" [pytest targets] [iters]
tag=$1; variants=$2; targets=${3:-tests/test_gpu_parity.py tests/test_gpu_golden_and_scale.py tests/test_gpu_full_length.py}; iters=${4:-50}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
# quick canary first: a wedged kernel must cost seconds, not the whole lease
timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stage_level and fp64 and True" > $out/${tag}_canary.log 2>&1; rc=$?
echo "canary rc=$rc: $(tail -1 $out/${tag}_canary.log)"
if [ $rc -ne 0 ]; then tail -30 $out/${tag}_canary.log; exit 1; fi
if [ "$targets" != none ]; then
  DCMA_PARITY_REPORT=$out/${tag}_parity.jsonl timeout 900 python -m pytest $targets -x -q -m gpu -s > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
  tail -5 $out/${tag}_pytest.log
fi
for v in main $variants; do
  if [ "$v" = main ]; then unset DCMA_B200_LIB; else export DCMA_B200_LIB=$PWD/dicomautomaton_b200/lib/libdcma_b200_$v.so; fi
  if [ "$v" != main ] && [ "$v" != notma ]; then
    timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stage_level or loop" > $out/${tag}_pytest_${v}.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $out/${tag}_pytest_${v}.log
    echo "== $v parity: $(tail -2 $out/${tag}_pytest_${v}.log | tr '\n' ' ')"
    if [ $rc -ne 0 ]; then echo "== $v skipped (parity failed or hung)"; continue; fi
  fi
  timeout 240 python bench.py --no-cpu-baseline --no-other-modes --steps 2 --warmup 3 --iters $iters --e2e-steps 1 > $out/${tag}_bench_${v}.json 2> $out/${tag}_bench_${v}.err
  echo "== $v"; python tools/show_bench.py $out/${tag}_bench_${v}.json | head -6
done
