#!/bin/bash
# last 1-GPU validation of the round: full -m gpu suite, smoke, the driver's default bench command
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c5_default.log 2>&1; echo "c5 (default flags) exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_c5_default.log'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value %.4g ms/step %.4f iters %.1f frac %.3f stored_gbs %.0f parity %s e2e %.4g steps %d warmup %d'%(d['value'],d['ms_per_step'],d['run_info']['cg_iters_per_step_mean'],r['frac'],r['stored_achieved_gbs'],d['parity']['ok'],d['e2e']['value'],d['steps'],d['warmup']), d['clocks'])
PY
