#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_c5.log','gpurun_out/bench_ref.log'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); e=d.get('e2e') or {}
            print(f, 'value %.4g ms/step %.2f e2e %.4g'%(d['value'], d['ms_per_step'], e.get('value') or 0), 'parity', (d.get('parity') or {}).get('ok'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
            print({k:v for k,v in e.items() if k!='api'})
PY
