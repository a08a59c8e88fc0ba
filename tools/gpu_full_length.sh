out=gpurun_out; mkdir -p $out
DCMA_PARITY_REPORT=$out/r2h_parity_main.jsonl timeout 600 python -m pytest tests/test_gpu_full_length.py -x -q -s -m gpu > $out/r2h_full_main.log 2>&1; echo "main rc=$?"; tail -5 $out/r2h_full_main.log
DCMA_B200_LIB=$PWD/dicomautomaton_b200/lib/libdcma_b200_emu.so DCMA_PARITY_REPORT=$out/r2h_parity_emu.jsonl timeout 600 python -m pytest tests/test_gpu_full_length.py -q -s -m gpu > $out/r2h_full_emu.log 2>&1; echo "emu rc=$?"; grep "fp64:" $out/r2h_full_emu.log
