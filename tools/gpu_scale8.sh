#!/bin/bash
# 8-GPU evidence (gpurun --gpus 8): NCCL slab parity, 1024^3 slab strong scaling point, independent registrations, config 5.
tag=$1; N=${2:-8}; out=gpurun_out; mkdir -p $out
export NCCL_DEBUG=WARN
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
nvidia-smi -L > $out/${tag}_smi.txt
timeout 300 $TR --master-port 29511 tests/multi_gpu/run_slab_nccl.py --shape 128,96,128 --iters 8 > $out/${tag}_slab_check.log 2>&1
echo "slab check rc=$?" | tee -a $out/${tag}_slab_check.log
timeout 600 $TR --master-port 29512 bench.py --slab --shape 1024,1024,1024 --gpus $N --steps 1 --warmup 1 --iters 10 > $out/${tag}_slab1024_n$N.json 2> $out/${tag}_slab1024_n$N.err
timeout 600 $TR --master-port 29513 bench.py --slab --gpus $N --steps 1 --warmup 1 --iters 30 > $out/${tag}_slab512_n$N.json 2> $out/${tag}_slab512_n$N.err
timeout 600 $TR --master-port 29514 bench.py --gpus $N --steps 1 --warmup 3 --iters 50 --no-cpu-baseline --no-other-modes > $out/${tag}_indep_n$N.json 2> $out/${tag}_indep_n$N.err
timeout 600 $TR --master-port 29515 tools/run_configs.py --configs 4,5 --out $out/${tag}_configs.json > $out/${tag}_configs.log 2>&1
python - $out $tag <<'PY'
import glob, json, sys
out, tag = sys.argv[1], sys.argv[2]
for f in sorted(glob.glob(f"{out}/{tag}_*_n*.json")):
    for l in open(f):
        if l.startswith('{'):
            d = json.loads(l)
            print(f.split('/')[-1], 'n_gpus', d['n_gpus'], 'value %.4g' % d['value'], 'ms/iter %.3f' % d['ms_per_iteration'], d.get('scaling'))
try:
    c = json.load(open(f"{out}/{tag}_configs.json"))
    for k, v in c["configs"].items():
        print("config", k, {kk: vv for kk, vv in v.items() if kk in ("loop_ms", "voxel_updates_per_s", "wall_s_max_over_ranks_incl_phantom_generation", "voxel_updates_per_s_loop_only", "ranks")})
except Exception as e:
    print("configs:", e)
PY
