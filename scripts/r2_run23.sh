#!/bin/bash
# slice metadata prefetched one iteration ahead in the TMA sweep: bitwise tests, then A/B on the same box against the previous
# build (TB_LIB), C5 and C2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_spmv_cg.py tests/test_gpu_c5_shape.py tests/test_gpu_monodomain.py tests/test_gpu_fusep.py tests/test_gpu_exact_dot.py tests/test_lv_config4.py tests/test_gpu_precond.py -m gpu -q --timeout=900 > gpurun_out/pytest_r2f.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_r2f.log
PREV=$PWD/thunderbolt.jl_b200/lib/libtbolt_b200_prev.so
for rep in 1 2; do
for v in prev new; do
if [ $v = prev ]; then export TB_LIB=$PREV; else unset TB_LIB; fi
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu --e2e-steps 0 --no-parity > gpurun_out/bench_c5_meta_$v.log 2>&1; echo "c5 $v exit $?"
grep '^{' gpurun_out/bench_c5_meta_$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  C5 $v: ms/step %.3f spmv %.4f ms stored %.0f GB/s'%(d['ms_per_step'], r['avg_launch_ms'], r['stored_achieved_gbs']))"
done
done
for v in prev new; do
if [ $v = prev ]; then export TB_LIB=$PREV; else unset TB_LIB; fi
timeout 300 python bench.py --workload c2 --steps 100 --warmup 20 --no-cpu --e2e-steps 0 --no-parity > gpurun_out/bench_c2_meta_$v.log 2>&1
timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu --e2e-steps 0 --no-parity > gpurun_out/bench_c4_meta_$v.log 2>&1
for w in c2 c4; do grep '^{' gpurun_out/bench_${w}_meta_$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print('  $w $v: ms/step %.4f its %.1f per-iteration %.4f ms'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], r['avg_launch_ms']))"; done
done
unset TB_LIB
