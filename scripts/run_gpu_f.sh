#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spmv_cg.py tests/test_gpu_monodomain.py tests/test_rtc.py tests/test_lv_config4.py -m gpu -q -x --timeout=600 > gpurun_out/pytest_cg.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_cg.log | cut -c1-300
for w in c1 c2; do
for pm in 1 0; do
TB_CG_PERSISTENT=$pm timeout 300 python bench.py --workload $w --steps 50 --warmup 3 --no-cpu > gpurun_out/bench_${w}_p$pm.log 2>&1
grep '^{' gpurun_out/bench_${w}_p$pm.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('$w persistent=$pm value %.4g ms/step %.3f iters %.1f unit_ms %.4f frac %.3f e2e %.4g'%(d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],r['avg_launch_ms'],r['frac'],d['e2e']['value']), r['kernel'][:30])"
grep -v '^{' gpurun_out/bench_${w}_p$pm.log | grep -iE "error|Traceback" -A5 | tail -8
done; done
