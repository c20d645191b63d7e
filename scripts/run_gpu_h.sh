#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_monodomain.py -m gpu -q -x --timeout=600 2>&1 | tail -3
for w in c5 c2 c1; do
timeout 400 python bench.py --workload $w --no-cpu > gpurun_out/bench_${w}_e2e.log 2>&1
grep '^{' gpurun_out/bench_${w}_e2e.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$w value %.4g ms/step %.3f e2e %.4g e2e ms/step %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))"
grep -v '^{' gpurun_out/bench_${w}_e2e.log | grep -iE "error|Traceback" -A5 | tail -8
done
