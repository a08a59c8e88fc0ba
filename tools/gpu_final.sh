#!/bin/bash
# Final evidence pass on one B200: tests, smoke, both bench arms, launch lists, ncu --set full of the loop kernels (fp64 headline + fp32).
tag=${1:-final}; out=gpurun_out; mkdir -p $out
bash tools/gpu_check.sh $tag
for mode in fp64 fp32; do
  python tools/profile_run.py --mode $mode --iters 2 > $out/${tag}_plain2_$mode.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'k_(warp_force|smooth3d|compose)' -s 4 -c 4 -f -o $out/${tag}_prof_$mode \
      python tools/profile_run.py --mode $mode --iters 2 > $out/${tag}_ncu2_$mode.log 2>&1
done
