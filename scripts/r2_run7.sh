#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_mesh_pattern.py tests/test_multidomain.py tests/test_lv_config4.py tests/test_ecg_leadfield.py tests/test_gpu_fusep.py -m gpu -q --timeout=1200 > gpurun_out/pytest_new2.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_new2.log
for f in 0 1; do
TB_SPMV_FUSEP=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/bench_c5_fusep$f.log 2>&1; echo "fusep$f exit $?"
done
python - <<'PY'
import json
for f in ('gpurun_out/bench_c5_fusep0.log','gpurun_out/bench_c5_fusep1.log'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, 'ms/step %.2f its %.1f spmv %.3f ms launches %d'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], d['roofline']['avg_launch_ms'], d['gpu_launches']))
PY
