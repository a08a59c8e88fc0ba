#!/bin/bash
# timing-only ablation variants (results are wrong by construction): prints the smoothing stage time per variant
out=gpurun_out; mkdir -p $out
for v in main $1; do
  if [ "$v" = main ]; then unset DCMA_B200_LIB; else export DCMA_B200_LIB=$PWD/dicomautomaton_b200/lib/libdcma_b200_$v.so; fi
  python bench.py --field-mode fp32 --no-cpu-baseline --steps 1 --warmup 3 --iters 30 --e2e-steps 1 > $out/abl_$v.json 2> $out/abl_$v.err
  echo "== $v"; python tools/show_bench.py $out/abl_$v.json | grep "stage ms"
done
