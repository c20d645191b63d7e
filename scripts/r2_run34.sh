#!/bin/bash
# tb_monodomain_run without read-backs (persistent CG paths): test, C1 / C2 lines with the run_api block
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_monodomain.py -m gpu -q --timeout=600 -k "run_without or random_inputs or fused_and_unfused" > gpurun_out/pytest_r2l.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_r2l.log
timeout 300 python bench.py --workload c1 --steps 1000 --warmup 100 > gpurun_out/bench_c1.log 2>&1; echo "c1 exit $?"
timeout 300 python bench.py --workload c2 --steps 100 --warmup 20 > gpurun_out/bench_c2.log 2>&1; echo "c2 exit $?"
python - <<'PY'
import json
for f in ['gpurun_out/bench_c1.log','gpurun_out/bench_c2.log']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, 'value %.4g ms/step %.4f its %.2f parity %s'%(d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], d['parity']['ok']), 'run_api', {k:v for k,v in (d.get('run_api') or {}).items() if k!='api'})
            break
    else: print(f, 'NO JSON', open(f).read()[-1500:])
PY
