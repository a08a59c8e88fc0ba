#!/bin/bash
# Scaling evidence on an N-GPU box (gpurun --gpus N): z-slab strong scaling at 512x512x256 and 1024^3 (BASELINE config 4),
# independent registrations (config 5 layout), NCCL parity check.   usage: bash tools/gpu_scale.sh <tag> <N> [mode]
tag=$1; N=$2; mode=${3:-fp64}; out=gpurun_out; mkdir -p $out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $out/${tag}_smi.txt
timeout 600 $TR --nproc-per-node $N --master-port 29511 tests/multi_gpu/run_slab_nccl.py --shape 96,96,128 --iters 8 > $out/${tag}_slab_check.log 2>&1
echo "slab check rc=$?" | tee -a $out/${tag}_slab_check.log
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  if [ $n -eq 1 ]; then L="python"; else L="$TR --nproc-per-node $n --master-port $((29520+n))"; fi
  timeout 900 $L bench.py --slab --field-mode $mode --gpus $n --steps 1 --warmup 1 --iters 30 > $out/${tag}_slab512_n$n.json 2> $out/${tag}_slab512_n$n.err
  timeout 1200 $L bench.py --slab --field-mode $mode --shape 1024,1024,1024 --gpus $n --steps 1 --warmup 1 --iters 10 > $out/${tag}_slab1024_n$n.json 2> $out/${tag}_slab1024_n$n.err
  timeout 900 $L bench.py --field-mode $mode --gpus $n --steps 1 --warmup 3 --iters 50 --no-cpu-baseline --no-other-modes > $out/${tag}_indep_n$n.json 2> $out/${tag}_indep_n$n.err
done
python - $out $tag <<'PY'
import glob, json, sys
out, tag = sys.argv[1], sys.argv[2]
for f in sorted(glob.glob(f"{out}/{tag}_*_n*.json")):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f.split('/')[-1], 'n_gpus', d['n_gpus'], 'value %.4g' % d['value'], 'ms/iter %.3f' % d['ms_per_iteration'], d.get('scaling'))
PY
