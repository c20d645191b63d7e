#!/bin/bash
# is fused != unfused on C4 (plain CG, 1154 iterations) new?  same check with the library as it was at the start of the session
mkdir -p gpurun_out
OLD=$PWD/thunderbolt.jl_b200/lib/libtbolt_b200_r2start.so
run() { name=$1; shift
env "$@" timeout 300 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/bench_c4_fu_$name.log 2>&1
grep '^{' gpurun_out/bench_c4_fu_$name.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); f=d['parity']['checks']['fused_vs_unfused']; print('  $name: ms/step %.2f its %.1f fused_vs_unfused %s'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], f))" || tail -5 gpurun_out/bench_c4_fu_$name.log
}
run old TB_LIB=$OLD
run new_hdr0 TB_SPMV_HDRONLY=0
run new
run new_pcg0 TB_CG_PERSISTENT=0
