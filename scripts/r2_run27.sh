#!/bin/bash
# from-b start of the persistent TMA CG sums like the fused start: apex-row test (bitwise fused == unfused), C4 parity block
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_lv_config4.py tests/test_gpu_spmv_cg.py tests/test_gpu_monodomain.py -m gpu -q --timeout=900 -k "not 1000_steps" > gpurun_out/pytest_r2h.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_r2h.log
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu --e2e-steps 0 > gpurun_out/bench_c4_plain.log 2>&1; echo "c4 exit $?"
grep '^{' gpurun_out/bench_c4_plain.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  C4 plain: ms/step %.3f its %.1f per-iteration %.4f ms parity %s'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], r['avg_launch_ms'], d['parity']['ok']), d['parity']['checks']['fused_vs_unfused'])"
