#!/bin/bash
# Round-2 multi-GPU pass (gpurun --gpus N): slab parity in fp64 / mixed / symmetric over the multi-process transport, the
# one-process multi-device path (incl. pyramid on slabs), and bench.py exactly as the driver launches it.
tag=$1; N=${2:-8}; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $out/${tag}_smi.txt
timeout 240 $TR --nproc-per-node $N --master-port 29511 tests/multi_gpu/run_slab_nccl.py --shape 256,256,256 --iters 10 --mode fp64 > $out/${tag}_check_fp64.log 2>&1; echo "slab check fp64 rc=$?" | tee -a $out/${tag}_check_fp64.log
timeout 240 $TR --nproc-per-node $N --master-port 29512 tests/multi_gpu/run_slab_nccl.py --shape 128,256,256 --iters 8 --mode mixed --symmetric > $out/${tag}_check_mixed_symmetric.log 2>&1; echo "slab check mixed+symmetric rc=$?" | tee -a $out/${tag}_check_mixed_symmetric.log
grep -h "rank 0/" $out/${tag}_check_fp64.log $out/${tag}_check_mixed_symmetric.log | cut -c1-260
timeout 300 python tests/multi_gpu/run_inprocess_devices.py --devices $N > $out/${tag}_inprocess.log 2>&1; echo "in-process rc=$?" | tee -a $out/${tag}_inprocess.log; grep "devices, one" $out/${tag}_inprocess.log
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d['config']['workload'][40:62], 'n',d['n_gpus'],'value %.4g'%d['value'],'ms/iter %.3f'%d['ms_per_iteration'],'xchg ms/iter %.3f'%d['exchange_ms_per_iter'], 'overlap',d['exchange_overlapped'], 'e2e %.3g'%d['e2e']['value'], 'check',d.get('slab_vs_whole_check'), 'replicas', (d.get('replicas') or {}).get('value'), d['clocks'])
PY
}
timeout 600 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > $out/${tag}_bench_driver.json 2> $out/${tag}_bench_driver.err; echo "driver-style rc=$?"
show $out/${tag}_bench_driver.json driver
if [ "${3:-}" = small ]; then
timeout 300 $TR --nproc-per-node $N --master-port 29524 bench.py --gpus $N --shape 256,512,512 --steps 3 --warmup 2 --no-replicas --no-slab-check --e2e-steps 1 > $out/${tag}_bench_small.json 2> $out/${tag}_bench_small.err
show $out/${tag}_bench_small.json small
fi
