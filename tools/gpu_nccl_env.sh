#!/bin/bash
# NCCL P2P channel settings vs slab step time (2 GPUs)
out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
i=0
for env in "A=1" "NCCL_MIN_P2P_NCHANNELS=8" "NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32" "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32" "NCCL_P2P_USE_CUDA_MEMCPY=1"; do
  i=$((i+1))
  env $env timeout 600 $TR --master-port $((29600+i)) bench.py --slab --gpus 2 --steps 1 --warmup 1 --iters 30 > $out/ncclenv_$i.json 2> $out/ncclenv_$i.err
  python - "$out/ncclenv_$i.json" "$env" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], '-> ms/iter %.3f'%d['ms_per_iteration'], 'value %.4g'%d['value'])
PY
done
