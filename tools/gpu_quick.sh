#!/bin/bash
# Quick GPU pass while iterating on kernels: parity tests, short bench in both modes, one ncu --set full capture.
# usage: bash tools/gpu_quick.sh <tag> [ncu-mode|none] [pytest -k expr]
tag=${1:-q}; ncumode=${2:-fp32}; kexpr=${3:-}
out=gpurun_out; mkdir -p $out
if [ -n "$kexpr" ]; then python -m pytest tests -x -q -m gpu -k "$kexpr" > $out/${tag}_pytest.log 2>&1; else python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; fi
echo "pytest rc=$?" >> $out/${tag}_pytest.log
for mode in fp32 fp64; do
  python bench.py --field-mode $mode --no-cpu-baseline --steps 2 --warmup 3 --iters 50 --e2e-steps 1 > $out/${tag}_bench_$mode.json 2> $out/${tag}_bench_$mode.err
done
if [ "$ncumode" != "none" ]; then
  python tools/profile_run.py --mode $ncumode --iters 2 > $out/${tag}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'k_(warp_force|smooth3d|compose)' -s 4 -c 4 -f -o $out/${tag}_prof \
      python tools/profile_run.py --mode $ncumode --iters 2 > $out/${tag}_ncu.log 2>&1
fi
tail -3 $out/${tag}_pytest.log; python tools/show_bench.py $out/${tag}_bench_fp32.json $out/${tag}_bench_fp64.json
