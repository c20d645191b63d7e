#!/bin/bash
# flag barrier with parallel polling (persistent CG variants 1 / 2 on C1), stream-ordered assembly, bench_assembly both element kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_spmv_cg.py tests/test_gpu_monodomain.py tests/test_gpu_assembly.py tests/test_gpu_assembly_gather.py tests/test_lv_config4.py tests/test_rtc.py -m gpu -q --timeout=900 > gpurun_out/pytest_r2c.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_r2c.log
for v in 1 2 1 2; do
TB_PCG_V=$v timeout 600 python bench.py --workload c1 --steps 1000 --warmup 100 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/bench_c1_pcgv$v.log 2>&1; echo "c1 v$v exit $?"
grep '^{' gpurun_out/bench_c1_pcgv$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('  C1 variant $v: ms/step %.4f its %.2f -> %.2f us/iteration'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], 1e3*d['ms_per_step']/d['run_info']['cg_iters_per_step_mean']))"
done
TB_ELEMENT_SPLIT=0 timeout 600 python scripts/bench_assembly.py --modes 2 > gpurun_out/bench_assembly_split0.log 2>&1; echo "asm split0 exit $?"
TB_ELEMENT_SPLIT=1 timeout 600 python scripts/bench_assembly.py --modes 2 > gpurun_out/bench_assembly_split1.log 2>&1; echo "asm split1 exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_assembly_split[01].log')):
    print(f)
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print('   %-22s %-18s %7.3f ms  %.3g el/s  frac %.3f'%(d['workload'], d['form'], d['ms'], d['elements_per_s'], d['frac']))
PY
