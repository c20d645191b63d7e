#!/bin/bash
# SpMV ring variants with the header-only stage (32 warps x 1 stage | 15 warps x 2 stages | 24 x 1 | 16 x 1), then the 1-GPU suite
mkdir -p gpurun_out
for v in 1 3 4 2; do
TB_SPMV_VARIANT=$v timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu --e2e-steps 0 --no-parity > gpurun_out/bench_c5_var$v.log 2>&1; echo "c5 variant $v exit $?"
grep '^{' gpurun_out/bench_c5_var$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  variant $v: ms/step %.3f spmv %.4f ms stored %.0f GB/s'%(d['ms_per_step'], r['avg_launch_ms'], r['stored_achieved_gbs']))"
done
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c5.log 2>&1; echo "c5 exit $?"
timeout 300 python bench.py --workload c2 --steps 100 --warmup 20 > gpurun_out/bench_c2.log 2>&1; echo "c2 exit $?"
timeout 600 python bench.py --workload c4 --precond block_jacobi --bj-rows 64 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c4_bj64.log 2>&1; echo "c4 exit $?"
python - <<'PY'
import json,glob
for f in ['gpurun_out/bench_c5.log','gpurun_out/bench_c2.log','gpurun_out/bench_c4_bj64.log']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(f, 'value %.4g ms/step %.4f iters %.1f frac %.3f stored_gbs %s parity %s'%(d['value'],d['ms_per_step'],d['run_info']['cg_iters_per_step_mean'],r['frac'],r.get('stored_achieved_gbs'),(d.get('parity') or {}).get('ok')), 'e2e %.4g'%((d.get('e2e') or {}).get('value') or 0), d['clocks'])
            break
    else: print(f,'NO JSON', open(f).read()[-800:])
PY
