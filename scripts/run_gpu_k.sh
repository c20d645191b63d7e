#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spmv_cg.py tests/test_gpu_monodomain.py -m gpu -q -x --timeout=600 2>&1 | tail -6 | cut -c1-300
for w in c2 c1; do
for pm in 1 2 0; do
TB_CG_PERSISTENT=$pm timeout 120 python bench.py --workload $w --steps 50 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/bench_${w}_p$pm.log 2>&1
grep '^{' gpurun_out/bench_${w}_p$pm.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('$w TB_CG_PERSISTENT=$pm value %.4g ms/step %.3f iters %.1f unit_ms %.4f frac %.3f e2e %.4g'%(d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],r['avg_launch_ms'],r['frac'],d['e2e']['value']), r['kernel'][:24])"
grep -v '^{' gpurun_out/bench_${w}_p$pm.log | grep -iE "error|Traceback" -A5 | tail -8
done; done
# mid-size 3D problems
for g in "256,256,64" "320,320,128"; do
for pm in 1 0; do
TB_CG_PERSISTENT=$pm timeout 120 python bench.py --workload c5 --grid $g --steps 10 --warmup 3 --no-cpu --e2e-steps 0 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('grid $g TB_CG_PERSISTENT=$pm dofs %d value %.4g ms/step %.3f iters %.1f unit_ms %.4f'%(d['config']['dofs'],d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],r['avg_launch_ms']), r['kernel'][:24])"
done; done
