#!/bin/bash
# Round-2 evidence pass on ONE B200: (1) the fp32-stored-U feasibility probe, (2) launch list + `ncu --set full` of the
# loop kernels of the shipped build (fp64 headline, fp32 beside it), (3) ncu of the kernels outside the loop.
# Every ncu command runs only after the same command exited 0 without ncu.  usage: bash tools/gpu_r2_ncu.sh [tag]
tag=${1:-r2n}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
if [ -f dicomautomaton_b200/lib/libdcma_b200_emu.so ]; then
DCMA_B200_LIB=$PWD/dicomautomaton_b200/lib/libdcma_b200_emu.so DCMA_PARITY_REPORT=$out/${tag}_parity_emu.jsonl timeout 400 python -m pytest tests/test_gpu_full_length.py -q -s -m gpu > $out/${tag}_full_emu.log 2>&1; echo "emu rc=$?"; grep "fp64:" $out/${tag}_full_emu.log
fi
for mode in fp64 fp32; do
  timeout 120 python tools/profile_run.py --mode $mode --iters 3 > $out/${tag}_plain_$mode.log 2>&1 || { echo "plain $mode failed"; continue; }
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_$mode.csv \
      python tools/profile_run.py --mode $mode --iters 3 > $out/${tag}_nculist_$mode.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_(warp_force|smooth3d|compose)' -s 4 -c 4 -f -o $out/${tag}_prof_$mode \
      python tools/profile_run.py --mode $mode --iters 2 > $out/${tag}_ncu_$mode.log 2>&1; echo "ncu $mode rc=$?"
done
timeout 300 python tools/profile_aux.py > $out/${tag}_aux_plain.log 2>&1; rc=$?; echo "aux plain rc=$rc"; tail -2 $out/${tag}_aux_plain.log
if [ $rc -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none -k regex:'^k_' -c 60 -f -o $out/${tag}_prof_aux python tools/profile_aux.py > $out/${tag}_ncu_aux.log 2>&1; echo "ncu aux rc=$?"
fi
ls -la $out | grep ${tag}_ | awk '{print $5, $9}'
