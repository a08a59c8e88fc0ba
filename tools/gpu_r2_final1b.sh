#!/bin/bash
# Final one-GPU record after the device pool: whole GPU suite, smoke, bench.py as the driver runs it.
tag=${1:-r2y}; out=gpurun_out; mkdir -p $out
DCMA_PARITY_REPORT=$out/${tag}_parity.jsonl timeout 1500 python -m pytest tests -x -q -m gpu ${PYTEST_ARGS:-} > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log; tail -6 $out/${tag}_pytest.log | cut -c1-300
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
DCMA_B200_TRACE=1 timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
grep "dcma_b200\] create" $out/${tag}_bench.err | tail -3
python - $out/${tag}_bench.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; e=d['e2e']
        print(' value %.4g ms/iter %.3f clocks %s launches %s' % (d['value'], d['ms_per_iteration'], d['clocks'], d['gpu_launches']))
        print(' dominant:', r['kernel'][:40], 'frac %.3f' % r['frac'], 'whole-iter frac %.3f' % r['whole_iteration']['frac_of_peak'])
        for k in r['per_kernel']: print('   %-28s %.3f ms  frac %.3f share %.3f' % (k['kernel'][:28], k['ms_per_launch'], k['frac'], k['share_of_step']))
        c=e['c_abi']; print(' e2e %.4g ms/step %.1f (ratio to C ABI %s) c_abi %.4g ms/step %.1f loop %.1f h2d %.1f d2h %.1f' % (e['value'], e['ms_per_step'], e.get('ratio_to_c_abi'), c['value'], c['ms_per_step'], c['loop_ms'], c['h2d_ms'], c['d2h_ms']))
        print(' cpu', d['cpu_baseline']['value'], ' other:', {k:(round(v['value_per_gpu']/1e9,2), round(v['ms_per_iteration'],3)) for k,v in d.get('other_modes',{}).items()})
PY
