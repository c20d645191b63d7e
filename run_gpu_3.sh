#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
TB_SPMV_VARIANT=$v timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu_v$v.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_v$v.log
done
for v in 0 1 2 3 4 5; do
TB_SPMV_VARIANT=$v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_c5_v$v.log 2>&1
done
for v in 0 1; do
TB_SPMV_VARIANT=$v timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c2_v$v.log 2>&1
done
tail -2 gpurun_out/pytest_gpu_v*.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*_v*.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(f, 'value %.4g ms/step %.2f iters %.1f spmv_ms %.3f frac %.3f share %.3f step_frac %.3f'%(d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],r['avg_launch_ms'],r['frac'],r['spmv_share_of_step'],r['step_frac_of_peak']))
            break
    else: print(f,'NO JSON', open(f).read()[-300:])
PY
