#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_c5_shape.py tests/test_gpu_exact_dot.py tests/test_gpu_precond.py tests/test_gpu_spmv_cg.py tests/test_rtc.py -m gpu -q --timeout=900 > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_new.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; echo "bench exit $?"
TB_DOT_EXACT=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_c5_exact.log 2>&1; echo "bench exact exit $?"
for pc in jacobi "block_jacobi --bj-rows 64" "block_jacobi --bj-rows 128" "chebyshev --cheb-degree 8 --cheb-ratio 100" "chebyshev --cheb-degree 16 --cheb-ratio 300"; do
  tag=$(echo $pc | tr ' -' '__')
  timeout 900 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --precond $pc > gpurun_out/bench_c4_$tag.log 2>&1; echo "c4 $pc exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*.log')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True
            e=d.get('e2e') or {}
            print(f, 'value %.4g ms/step %.2f its %.1f e2e %s parity %s'%(d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], e.get('value'), (d.get('parity') or {}).get('ok')))
            if d.get('parity') and not d['parity']['ok']: print(json.dumps(d['parity'])[:1500])
    if not ok: print(f, 'NO JSON', open(f).read()[-1200:])
PY
