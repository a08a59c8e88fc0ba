#!/bin/bash
# Round-2 one-GPU pass E: new entry points (register_stacks, virtual fields), the drop-in on the one-call path, e2e.
tag=${1:-r2e2}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_extensions.py tests/test_host_dropin.py tests/test_operation_dropin.py tests/test_warp_operation.py -q -m gpu > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log; tail -15 $out/${tag}_pytest.log | cut -c1-300
DCMA_B200_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --iters 50 --no-cpu-baseline --no-other-modes --dropin-steps 3 --e2e-steps 2 > $out/${tag}_trace_bench.json 2> $out/${tag}_trace_bench.err; echo "trace bench rc=$?"; tail -8 $out/${tag}_trace_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-modes > $out/${tag}_bench_fp64.json 2> $out/${tag}_bench_fp64.err; echo "bench rc=$?"
python - $out/${tag}_bench_fp64.json <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print(' value %.4g ms/iter %.3f clocks %s' % (d['value'], d['ms_per_iteration'], d['clocks']))
        print(' e2e %.4g ms/step %.1f (ratio to C ABI %s) c_abi %.4g ms/step %.1f loop %.1f' % (e['value'], e['ms_per_step'], e.get('ratio_to_c_abi'), e['c_abi']['value'], e['c_abi']['ms_per_step'], e['c_abi']['loop_ms']))
PY
