#!/bin/bash
# bench.py exactly as the driver launches it on N GPUs, nothing else.  usage: bash tools/gpu_r2_bench8.sh <tag> [N]
tag=$1; N=${2:-8}; out=gpurun_out; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29521 bench.py --gpus $N --steps 5 --warmup 3 > $out/${tag}_bench_driver.json 2> $out/${tag}_bench_driver.err; echo "driver-style rc=$?"
python - $out/${tag}_bench_driver.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('n',d['n_gpus'],'value %.4g'%d['value'],'ms/iter %.3f'%d['ms_per_iteration'],'e2e %.4g ms/step %.1f h2d %.2f GB d2h %.2f GB nvlink %.2f GB'%(e['value'],e['ms_per_step'],e['h2d_bytes_per_step']/1e9,e['d2h_bytes_per_step']/1e9,e.get('nvlink_bytes_per_step',0)/1e9),'check',(d.get('slab_vs_whole_check') or {}).get('bit_identical_on_all_ranks'),'replicas %.4g'%(d.get('replicas') or {}).get('value',0), d['clocks'])
PY
