#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=700 -k "2]" > gpurun_out/pytest_multi2.log 2>&1; echo "multi exit $?"; tail -n 12 gpurun_out/pytest_multi2.log
timeout 900 python -m pytest tests/test_gpu_precond.py tests/test_gpu_exact_dot.py -m gpu -q --timeout=700 > gpurun_out/pytest_pc.log 2>&1; echo "pc exit $?"; tail -n 5 gpurun_out/pytest_pc.log
for r in 64 96; do
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --precond block_jacobi --bj-rows $r > gpurun_out/bench_c4_bj$r.log 2>&1; echo "c4 bj$r exit $?"
done
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --precond chebyshev --cheb-degree 8 --cheb-ratio 100 > gpurun_out/bench_c4_cheb8.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c4_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            print(f, 'N=%d value %.4g ms/step %.2f its %.1f parity %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], (d.get('parity') or {}).get('ok')))
PY
