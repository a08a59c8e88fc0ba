#!/bin/bash
# 8-GPU pass (gpurun --gpus 8): slab parity, bench.py exactly as the driver launches it, and the exchange ablations.
tag=$1; N=${2:-8}; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $out/${tag}_smi.txt; free -g > $out/${tag}_mem.txt
timeout 240 $TR --nproc-per-node $N --master-port 29511 tests/multi_gpu/run_slab_nccl.py --shape 256,256,256 --iters 10 --mode fp64 > $out/${tag}_check.log 2>&1
echo "slab check fp64 rc=$?" | tee -a $out/${tag}_check.log
grep "rank" $out/${tag}_check.log | grep "diffeo=True" | head -8
show() { python - "$1" "$2" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[2], d['config']['workload'][40:62], 'n',d['n_gpus'],'value %.4g'%d['value'],'ms/iter %.3f'%d['ms_per_iteration'],'xchg ms/iter %.3f'%d['exchange_ms_per_iter'], 'overlap',d['exchange_overlapped'], 'e2e %.3g'%d['e2e']['value'], 'check',d.get('slab_vs_whole_check'), 'replicas', (d.get('replicas') or {}).get('value'), d['clocks'])
PY
}
timeout 600 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --steps 10 --warmup 2 > $out/${tag}_bench_driver.json 2> $out/${tag}_bench_driver.err; echo "driver-style rc=$?"
show $out/${tag}_bench_driver.json driver
export DCMA_B200_SLAB_NO_OVERLAP=1
timeout 300 $TR --nproc-per-node $N --master-port 29522 bench.py --gpus $N --steps 3 --warmup 1 --no-replicas --no-slab-check --e2e-steps 1 > $out/${tag}_bench_nooverlap.json 2> $out/${tag}_bench_nooverlap.err
show $out/${tag}_bench_nooverlap.json nooverlap
timeout 300 $TR --nproc-per-node $N --master-port 29523 bench.py --gpus $N --shape 256,512,512 --steps 3 --warmup 1 --no-replicas --no-slab-check --e2e-steps 1 > $out/${tag}_bench_nooverlap_small.json 2> $out/${tag}_bench_nooverlap_small.err
show $out/${tag}_bench_nooverlap_small.json nooverlap
unset DCMA_B200_SLAB_NO_OVERLAP
timeout 300 $TR --nproc-per-node $N --master-port 29524 bench.py --gpus $N --shape 256,512,512 --steps 3 --warmup 1 --no-replicas --no-slab-check --e2e-steps 1 > $out/${tag}_bench_small.json 2> $out/${tag}_bench_small.err
show $out/${tag}_bench_small.json default
