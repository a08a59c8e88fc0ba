#!/bin/bash
tag=$1; N=${2:-2}; out=gpurun_out; mkdir -p $out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 300 $TR --master-port 29511 tests/multi_gpu/run_slab_nccl.py --shape 64,96,128 --iters 10 > $out/${tag}_slab_check.log 2>&1
echo "slab check rc=$?"; grep rank $out/${tag}_slab_check.log | head -8; tail -3 $out/${tag}_slab_check.log | grep -v rank
for ipc in 0 1; do
  DCMA_B200_SLAB_NO_IPC=$ipc timeout 600 $TR --master-port $((29520+ipc)) bench.py --slab --gpus $N --steps 1 --warmup 1 --iters 30 > $out/${tag}_slab512_noipc$ipc.json 2> $out/${tag}_slab512_noipc$ipc.err
  DCMA_B200_SLAB_NO_IPC=$ipc timeout 600 $TR --master-port $((29530+ipc)) bench.py --slab --shape 1024,1024,1024 --gpus $N --steps 1 --warmup 1 --iters 10 > $out/${tag}_slab1024_noipc$ipc.json 2> $out/${tag}_slab1024_noipc$ipc.err
done
python - $out $tag <<'PY'
import glob, json, sys
out, tag = sys.argv[1], sys.argv[2]
for f in sorted(glob.glob(f"{out}/{tag}_slab*_noipc*.json")):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f.split('/')[-1], 'n_gpus', d['n_gpus'], 'value %.4g' % d['value'], 'ms/iter %.3f' % d['ms_per_iteration'])
PY
