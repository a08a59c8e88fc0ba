#!/bin/bash
# Final one-GPU record of round 2: the whole GPU suite, smoke, bench.py exactly as the driver runs it (both arms), and the
# ncu launch list of the loop kernels of the bench command.
tag=${1:-r2z}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
DCMA_PARITY_REPORT=$out/${tag}_parity.jsonl timeout 1500 python -m pytest tests -x -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log; tail -6 $out/${tag}_pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; echo "ref rc=$?"; cut -c1-200 $out/${tag}_bench_reference.json
timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - $out/${tag}_bench.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; e=d['e2e']
        print(' value %.4g ms/iter %.3f clocks %s launches %s' % (d['value'], d['ms_per_iteration'], d['clocks'], d['gpu_launches']))
        print(' dominant:', r['kernel'][:40], 'frac %.3f' % r['frac'], 'whole-iter frac %.3f' % r['whole_iteration']['frac_of_peak'], 'traffic', r['traffic'])
        for k in r['per_kernel']: print('   %-28s %.3f ms  frac %.3f share %.3f' % (k['kernel'][:28], k['ms_per_launch'], k['frac'], k['share_of_step']))
        print(' e2e %.4g ms/step %.1f (ratio to C ABI %s) c_abi %.4g' % (e['value'], e['ms_per_step'], e.get('ratio_to_c_abi'), e['c_abi']['value']))
        print(' cpu', d['cpu_baseline'])
        print(' other:', {k:(round(v['value_per_gpu']/1e9,2), round(v['ms_per_iteration'],3)) for k,v in d.get('other_modes',{}).items()})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_' -c 200 --csv --log-file $out/${tag}_launches_bench.csv python bench.py --steps 1 --warmup 3 --iters 10 --no-cpu-baseline --no-other-modes --no-dropin --e2e-steps 1 > $out/${tag}_nculist.log 2>&1; echo "ncu list rc=$?"
