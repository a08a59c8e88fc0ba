#!/bin/bash
# ncu --set full of the loop kernels for one library variant and mode: bash tools/gpu_ncu_variant.sh <tag> <variant|main> <mode>
tag=$1; v=$2; mode=$3; out=gpurun_out; mkdir -p $out
if [ "$v" != main ]; then export DCMA_B200_LIB=$PWD/dicomautomaton_b200/lib/libdcma_b200_$v.so; fi
python tools/profile_run.py --mode $mode --iters 2 > $out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_(warp_force|smooth3d|compose)' -s 4 -c 4 -f -o $out/${tag}_prof \
    python tools/profile_run.py --mode $mode --iters 2 > $out/${tag}_ncu.log 2>&1
tail -2 $out/${tag}_ncu.log
