#!/bin/bash
# GPU pass for the field-text calls (row f3): plain-C caller, C++ host writer/reader, then the pytest file.
out=gpurun_out; mkdir -p $out
timeout 40 tests/native/_build/text_io_gpu_check ${1:-12000000} > $out/r1_text_c.log 2>&1; echo "rc=$?" >> $out/r1_text_c.log
timeout 15 tests/native/_build/field_io_check > $out/r1_text_cc.log 2>&1; echo "rc=$?" >> $out/r1_text_cc.log
cat $out/r1_text_c.log $out/r1_text_cc.log
timeout 120 python -m pytest tests/test_field_text.py -x -q -m gpu > $out/r1_text_pytest.log 2>&1; echo "rc=$?" >> $out/r1_text_pytest.log
tail -5 $out/r1_text_pytest.log
