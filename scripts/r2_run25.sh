#!/bin/bash
# wide (apex) rows with eight chunks in flight: LV tests, C4 without / with block-Jacobi
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_lv_config4.py tests/test_gpu_spmv_cg.py tests/test_gpu_precond.py -m gpu -q --timeout=900 > gpurun_out/pytest_r2g.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_r2g.log
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu --e2e-steps 0 > gpurun_out/bench_c4_plain.log 2>&1; echo "c4 exit $?"
timeout 600 python bench.py --workload c4 --precond block_jacobi --bj-rows 64 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c4_bj64.log 2>&1; echo "c4 bj exit $?"
timeout 600 python bench.py --workload c4 --precond jacobi --steps 3 --warmup 2 --no-cpu --e2e-steps 0 --no-parity > gpurun_out/bench_c4_jacobi.log 2>&1; echo "c4 jacobi exit $?"
for f in plain bj64 jacobi; do grep '^{' gpurun_out/bench_c4_$f.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  C4 $f: ms/step %.3f its %.1f per-iteration %.4f ms value %.4g parity %s'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], r['avg_launch_ms'], d['value'], (d.get('parity') or {}).get('ok')))"; done
