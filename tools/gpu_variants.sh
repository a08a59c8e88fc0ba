#!/bin/bash
# Bench tuning variants of the library (built by tools/build_variant.py) in one GPU call.
# usage: bash tools/gpu_variants.sh <tag> "<variant tags>" [modes]
tag=$1; variants=$2; modes=${3:-fp32}
out=gpurun_out; mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log; tail -2 $out/${tag}_pytest.log
[ -x tools/ubench/ffma2 ] && tools/ubench/ffma2 > $out/${tag}_ffma2.txt 2>&1 && cat $out/${tag}_ffma2.txt
for v in main $variants; do
  for mode in $modes; do
    if [ "$v" = main ]; then unset DCMA_B200_LIB; else export DCMA_B200_LIB=$PWD/dicomautomaton_b200/lib/libdcma_b200_$v.so; fi
    if [ "$v" != main ] && [ "$mode" = "${modes%% *}" ]; then  # parity of the variant (stage level + loops), once per variant
      python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $out/${tag}_pytest_${v}.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_${v}.log
      echo "== $v parity: $(tail -2 $out/${tag}_pytest_${v}.log | tr '\n' ' ')"
    fi
    python bench.py --field-mode $mode --no-cpu-baseline --steps 2 --warmup 3 --iters 50 --e2e-steps 1 > $out/${tag}_bench_${v}_$mode.json 2> $out/${tag}_bench_${v}_$mode.err
    echo "== $v $mode"; python tools/show_bench.py $out/${tag}_bench_${v}_$mode.json | head -4
  done
done
