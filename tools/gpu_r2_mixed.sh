#!/bin/bash
# Round-2 one-GPU pass for the mixed field mode (D fp64 + U fp32) and the host-side changes: whole GPU suite, smoke, the
# bench in fp64 (as the driver runs it) and in mixed, ncu of the kernels outside the loop.  usage: bash tools/gpu_r2_mixed.sh [tag]
tag=${1:-r2m}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stage_level and mixed and True" > $out/${tag}_canary.log 2>&1; rc=$?
echo "canary rc=$rc: $(tail -1 $out/${tag}_canary.log)"; if [ $rc -ne 0 ]; then tail -30 $out/${tag}_canary.log; fi
DCMA_PARITY_REPORT=$out/${tag}_parity.jsonl timeout 1200 python -m pytest tests -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log; tail -15 $out/${tag}_pytest.log | cut -c1-300
cat $out/${tag}_parity.jsonl 2>/dev/null | cut -c1-600
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke\]" $out/${tag}_smoke.log
for mode in mixed fp64; do
  timeout 600 python bench.py --field-mode $mode --steps 3 --warmup 3 --no-cpu-baseline --dropin-steps 4 > $out/${tag}_bench_$mode.json 2> $out/${tag}_bench_$mode.err; echo "bench $mode rc=$?"
  python - $out/${tag}_bench_$mode.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; e=d['e2e']
        print(' value %.4g ms/iter %.3f clocks %s' % (d['value'], d['ms_per_iteration'], d['clocks']))
        print(' dominant:', r['kernel'][:40], 'frac %.3f' % r['frac'], 'whole-iter frac %.3f' % r['whole_iteration']['frac_of_peak'])
        for k in r['per_kernel']: print('   %-28s %.3f ms  frac %.3f share %.3f' % (k['kernel'][:28], k['ms_per_launch'], k['frac'], k['share_of_step']))
        print(' e2e %.4g (ratio to C ABI %s) c_abi %.4g' % (e['value'], e.get('ratio_to_c_abi'), (e.get('c_abi') or e)['value']))
        print(' other:', {k:(round(v['value_per_gpu']/1e9,2), round(v['ms_per_iteration'],3)) for k,v in d.get('other_modes',{}).items()})
PY
done
timeout 300 python tools/profile_aux.py > $out/${tag}_aux_plain.log 2>&1; rc=$?; echo "aux plain rc=$rc"; tail -2 $out/${tag}_aux_plain.log
if [ $rc -eq 0 ]; then
  timeout 900 ncu --set full --clock-control none -k regex:'^k_' -c 60 -f -o $out/${tag}_prof_aux python tools/profile_aux.py > $out/${tag}_ncu_aux.log 2>&1; echo "ncu aux rc=$?"
fi
timeout 120 python tools/profile_run.py --mode mixed --iters 2 > $out/${tag}_plain_mixed.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_(warp_force|smooth3d|compose)' -s 4 -c 4 -f -o $out/${tag}_prof_mixed \
      python tools/profile_run.py --mode mixed --iters 2 > $out/${tag}_ncu_mixed.log 2>&1; echo "ncu mixed rc=$?"
