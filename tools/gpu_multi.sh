#!/bin/bash
# Multi-GPU pass (run under `gpurun --gpus N`): NCCL slab parity check, slab strong scaling, independent-registration scaling.
# usage: bash tools/gpu_multi.sh <tag> <N>
tag=$1; N=$2; out=gpurun_out; mkdir -p $out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $out/${tag}_smi.txt
timeout 600 $TR --nproc-per-node $N --master-port 29511 tests/multi_gpu/run_slab_nccl.py --shape 64,96,128 --iters 10 > $out/${tag}_slab_check.log 2>&1
echo "slab check rc=$?" | tee -a $out/${tag}_slab_check.log
timeout 600 $TR --nproc-per-node $N --master-port 29512 tests/multi_gpu/run_slab_nccl.py --shape 96,128,160 --iters 12 --mode fp32 >> $out/${tag}_slab_check.log 2>&1
echo "slab check fp32 rc=$?" | tee -a $out/${tag}_slab_check.log
grep "rank" $out/${tag}_slab_check.log | head -20
python bench.py --slab --gpus 1 --steps 1 --warmup 1 --iters 30 > $out/${tag}_slab_n1.json 2> $out/${tag}_slab_n1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 $TR --nproc-per-node $n --master-port $((29520+n)) bench.py --slab --gpus $n --steps 1 --warmup 1 --iters 30 > $out/${tag}_slab_n$n.json 2> $out/${tag}_slab_n$n.err
    timeout 900 $TR --nproc-per-node $n --master-port $((29530+n)) bench.py --gpus $n --steps 1 --warmup 3 --iters 50 --no-cpu-baseline --no-other-modes > $out/${tag}_indep_n$n.json 2> $out/${tag}_indep_n$n.err
  fi
done
for f in $out/${tag}_slab_n*.json $out/${tag}_indep_n*.json; do python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1], 'n_gpus',d['n_gpus'],'value %.4g'%d['value'],'ms/iter %.3f'%d['ms_per_iteration'], d.get('scaling'))
PY
done
