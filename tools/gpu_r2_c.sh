#!/bin/bash
# Round-2 one-GPU pass C: changed kernels (mixed compose, warp_to_grid), drop-in end-to-end after the helper-thread change,
# config-3 quality on the multiscale phantom, per-GPU compute time of one 128-slice slab of the 1024^3 workload.
tag=${1:-r2c}; out=gpurun_out; mkdir -p $out
DCMA_PARITY_REPORT=$out/${tag}_parity.jsonl timeout 900 python -m pytest tests/test_gpu_extensions.py tests/test_gpu_full_length.py tests/test_host_dropin.py tests/test_gpu_slab.py tests/test_field_text.py tests/test_gpu_parity.py -q -m gpu -k "not fp32" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log; tail -12 $out/${tag}_pytest.log | cut -c1-300
grep -o '"mixed": {[^}]*}' $out/${tag}_parity.jsonl
for mode in fp64 mixed; do
  timeout 600 python bench.py --field-mode $mode --steps 3 --warmup 3 --no-cpu-baseline --no-other-modes --dropin-steps 6 > $out/${tag}_bench_$mode.json 2> $out/${tag}_bench_$mode.err; echo "bench $mode rc=$?"
  python - $out/${tag}_bench_$mode.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; e=d['e2e']
        print(' value %.4g ms/iter %.3f clocks %s' % (d['value'], d['ms_per_iteration'], d['clocks']))
        for k in r['per_kernel']: print('   %-28s %.3f ms  frac %.3f share %.3f' % (k['kernel'][:28], k['ms_per_launch'], k['frac'], k['share_of_step']))
        print(' e2e %.4g ms/step %.1f (ratio to C ABI %s) c_abi %.4g ms/step %.1f loop %.1f h2d %.1f d2h %.1f' % (e['value'], e['ms_per_step'], e.get('ratio_to_c_abi'), e['c_abi']['value'], e['c_abi']['ms_per_step'], e['c_abi']['loop_ms'], e['c_abi']['h2d_ms'], e['c_abi']['d2h_ms']))
PY
done
timeout 600 python tools/config3_eval.py --shape 512,512,512 --texture multiscale --out $out/${tag}_config3_multiscale.json > $out/${tag}_config3.log 2>&1; echo "config3 rc=$?"; cut -c1-330 $out/${tag}_config3.log
timeout 400 python bench.py --slab --gpus 1 --shape 128,1024,1024 --steps 2 --warmup 1 --slab-iters 20 --e2e-steps 1 --no-slab-check > $out/${tag}_slab_one_slab.json 2> $out/${tag}_slab_one_slab.err; echo "one-slab rc=$?"
python - $out/${tag}_slab_one_slab.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('one 128-slice slab of 1024^2: ms/iter %.3f value %.4g clocks %s' % (d['ms_per_iteration'], d['value'], d['clocks']))
PY
