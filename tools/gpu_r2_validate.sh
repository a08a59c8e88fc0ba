#!/bin/bash
# One-GPU validation pass of round 2: the whole GPU suite, the bench as the driver runs it, the 1-GPU baseline of the sharded
# workload, the config-3 quality report and the fp32-U feasibility probe.
tag=${1:-r2v}; out=gpurun_out; mkdir -p $out
DCMA_PARITY_REPORT=$out/${tag}_parity.jsonl timeout 1100 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log; tail -4 $out/${tag}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; python tools/show_bench.py $out/${tag}_bench.json | head -8
python - $out/${tag}_bench.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('e2e', {k:e.get(k) for k in ('value','ms_per_step','steps','ratio_to_c_abi','agrees_with_c_abi_at_probe_voxel')}, 'c_abi', (e.get('c_abi') or {}).get('value'))
PY
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $out/${tag}_bench_ref.json
timeout 500 python bench.py --slab --gpus 1 --steps 2 --warmup 1 --slab-iters 20 --e2e-steps 1 > $out/${tag}_slab_n1.json 2> $out/${tag}_slab_n1.err; echo "slab n1 rc=$?"
python - $out/${tag}_slab_n1.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('slab N=1', d['config']['workload'][40:70], 'value %.4g'%d['value'], 'ms/iter %.3f'%d['ms_per_iteration'], 'e2e %.3g'%d['e2e']['value'], d['clocks'])
PY
timeout 600 python tools/config3_eval.py --shape 512,512,512 --out $out/${tag}_config3.json > $out/${tag}_config3.log 2>&1; echo "config3 rc=$?"; cat $out/${tag}_config3.log | cut -c1-400
DCMA_B200_LIB=$PWD/dicomautomaton_b200/lib/libdcma_b200_emu.so DCMA_PARITY_REPORT=$out/${tag}_parity_emu.jsonl timeout 400 python -m pytest tests/test_gpu_full_length.py -q -s -m gpu > $out/${tag}_full_emu.log 2>&1; echo "emu rc=$?"; grep "fp64:" $out/${tag}_full_emu.log
