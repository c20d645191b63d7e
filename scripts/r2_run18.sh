#!/bin/bash
# ticket barrier in the two-barrier persistent CG (C1, variants 1 / 2); half-traffic block-Jacobi apply (C4, TB_BJ_SYM 0 / 1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_spmv_cg.py tests/test_gpu_precond.py tests/test_gpu_monodomain.py tests/test_rtc.py tests/test_multidomain.py -m gpu -q --timeout=900 > gpurun_out/pytest_r2d.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_r2d.log
for v in 1 2 1 2; do
TB_PCG_V=$v timeout 600 python bench.py --workload c1 --steps 1000 --warmup 100 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/bench_c1_pcgv$v.log 2>&1; echo "c1 v$v exit $?"
grep '^{' gpurun_out/bench_c1_pcgv$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  C1 variant $v: ms/step %.4f its %.2f -> %.2f us/iteration'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], 1e3*d['ms_per_step']/d['run_info']['cg_iters_per_step_mean']))"
done
for v in 0 1; do
TB_BJ_SYM=$v timeout 600 python bench.py --workload c4 --precond block_jacobi --bj-rows 64 --steps 5 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_c4_bj64_sym$v.log 2>&1; echo "c4 sym$v exit $?"
grep '^{' gpurun_out/bench_c4_bj64_sym$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  C4 bj64 sym=$v: ms/step %.3f its %.1f value %.4g parity %s'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], d['value'], (d.get('parity') or {}).get('ok')))"
done
TB_BJ_SYM=1 timeout 600 python bench.py --workload c4 --precond block_jacobi --bj-rows 96 --steps 5 --warmup 3 --no-cpu --e2e-steps 0 --no-parity > gpurun_out/bench_c4_bj96_sym1.log 2>&1
grep '^{' gpurun_out/bench_c4_bj96_sym1.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  C4 bj96 sym=1: ms/step %.3f its %.1f value %.4g'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], d['value']))"
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:k_bj_apply -s 20 -c 2 -o gpurun_out/prof_bj_apply_sym python bench.py --workload c4 --precond block_jacobi --bj-rows 64 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/ncu_bj_sym.log 2>&1; echo "ncu bj exit $?"
