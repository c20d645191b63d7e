#!/bin/bash
# round 2, call 1: new parity tests, full gpu suite, C5 bench with parity block (+ gpu_n1 golden), reference arm timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_c5_shape.py -m gpu -q -x --timeout=800 > gpurun_out/pytest_c5shape.log 2>&1; echo "c5shape exit $?"; tail -n 5 gpurun_out/pytest_c5shape.log
timeout 600 python bench.py --steps 5 --warmup 3 --write-golden > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_c5.err
cp tests/golden/c5_checksum.json gpurun_out/c5_checksum.json
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"; tail -n 4 gpurun_out/bench_ref.log | cut -c1-400
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "c2 exit $?"
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --deselect tests/test_gpu_c5_shape.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ('gpurun_out/bench_c5.log','gpurun_out/bench_c2.log'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, 'value %.4g ms/step %.2f e2e %.4g'%(d['value'], d['ms_per_step'], d['e2e']['value']))
            print(json.dumps(d['parity'])[:3000])
PY
