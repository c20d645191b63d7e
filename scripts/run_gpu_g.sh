#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmv_cg.py tests/test_rtc.py -m gpu -q -x --timeout=600 2>&1 | tail -3
for v in 1 3 2 4; do
for w in c2 c5; do
steps=5; [ $w = c2 ] && steps=30
TB_SPMV_VARIANT=$v timeout 300 python bench.py --workload $w --steps $steps --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_${w}_v$v.log 2>&1
grep '^{' gpurun_out/bench_${w}_v$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('$w variant=$v value %.4g ms/step %.3f iters %.1f spmv_ms %.4f frac %.3f stored_gbs %.0f'%(d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],r['avg_launch_ms'],r['frac'],r['stored_achieved_gbs']))"
grep -v '^{' gpurun_out/bench_${w}_v$v.log | grep -iE "error|Traceback" -A5 | tail -8
done; done
