#!/bin/bash
# Slab pass on N GPUs (gpurun --gpus N): parity of the sharded path against the whole-volume engine, then bench.py the way
# the driver launches it.   usage: bash tools/gpu_slab2.sh <tag> <N> [bench args]
tag=$1; N=$2; shift 2; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $out/${tag}_smi.txt
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
timeout 300 $TR --nproc-per-node $N --master-port 29511 tests/multi_gpu/run_slab_nccl.py --shape 64,96,128 --iters 10 > $out/${tag}_check.log 2>&1
echo "slab check strict rc=$?" | tee -a $out/${tag}_check.log
timeout 300 $TR --nproc-per-node $N --master-port 29512 tests/multi_gpu/run_slab_nccl.py --shape 128,256,256 --iters 12 --mode fp64 >> $out/${tag}_check.log 2>&1
echo "slab check fp64 rc=$?" | tee -a $out/${tag}_check.log
grep "rank\|rc=" $out/${tag}_check.log | head -24
for v in default nooverlap tokens; do
  unset DCMA_B200_SLAB_NO_OVERLAP DCMA_B200_SLAB_NCCL_TOKENS
  [ $v = nooverlap ] && export DCMA_B200_SLAB_NO_OVERLAP=1
  [ $v = tokens ] && export DCMA_B200_SLAB_NCCL_TOKENS=1
  for shape in 256,512,512 1024,1024,1024; do
    timeout 600 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --shape $shape --no-replicas --no-slab-check --e2e-steps 1 "$@" > $out/${tag}_bench_${v}_${shape//,/x}.json 2> $out/${tag}_bench_${v}_${shape//,/x}.err
    python - $out/${tag}_bench_${v}_${shape//,/x}.json $v <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d['config']['workload'][:60], 'n',d['n_gpus'],'value %.4g'%d['value'],'ms/iter %.3f'%d['ms_per_iteration'],'xchg ms/iter %.3f'%d['exchange_ms_per_iter'], 'overlap',d['exchange_overlapped'], d['transport'][:12], 'e2e %.3g'%d['e2e']['value'], d['clocks'])
PY
  done
done
