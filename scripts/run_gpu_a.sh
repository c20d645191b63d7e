#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_assembly_gather.py tests/test_gpu_assembly.py tests/test_lv_config4.py -m gpu -q -x --timeout=300 > gpurun_out/pytest_asm.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_asm.log
tail -15 gpurun_out/pytest_asm.log
timeout 600 python scripts/bench_assembly.py > gpurun_out/bench_assembly.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/bench_assembly.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['workload'], d['form'], 'mode', d['mode_requested'], d['mode_used'], 'chunks', d['chunks'], 'ms %.2f'%d['ms'], 'Mel/s %.1f'%(d['elements_per_s']/1e6), 'GB/s %.0f frac %.3f'%(d['achieved_gbs'], d['frac']))
    else: print(l.rstrip()[-300:])
PY
