#!/bin/bash
# ticket barrier in the mid-size persistent CG kernel (k_cg_persistent_tma): tests, C2 / C4 with TB_PCG_V = 1 (grid.sync) / 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_spmv_cg.py tests/test_gpu_monodomain.py tests/test_lv_config4.py tests/test_gpu_precond.py tests/test_rtc.py -m gpu -q --timeout=900 -k "not 1000_steps" > gpurun_out/pytest_r2m.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_r2m.log
for v in 1 2 1 2; do
TB_PCG_V=$v timeout 300 python bench.py --workload c2 --steps 200 --warmup 20 --no-cpu --e2e-steps 0 --no-parity > gpurun_out/bench_c2_pcgv$v.log 2>&1
grep '^{' gpurun_out/bench_c2_pcgv$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  C2 variant $v: ms/step %.4f its %.1f per-iteration %.2f us'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], 1e3*d['roofline']['avg_launch_ms']))"
done
for v in 1 2; do
TB_PCG_V=$v timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu --e2e-steps 0 > gpurun_out/bench_c4_pcgv$v.log 2>&1
grep '^{' gpurun_out/bench_c4_pcgv$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  C4 variant $v: ms/step %.3f its %.1f per-iteration %.2f us parity %s'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], 1e3*d['roofline']['avg_launch_ms'], d['parity']['ok']))"
done
