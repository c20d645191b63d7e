#!/bin/bash
# 4 GPUs on the final code (the slot-table gather is active from N = 4 on C5): the driver's scaling command, then the 2- and
# 4-rank distributed checks
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 4 --master-port 29664 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_c5_n4.log 2>&1; echo "c5 n4 exit $?"
python - <<'PY'
import json
f='gpurun_out/bench_c5_n4.log'; ok=False
for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); ok=True; e=d.get('e2e') or {}
        print('N=%d value %.4g ms/step %.2f its %s e2e %s parity %s setup %.1fs assembly: %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], e.get('value'), (d.get('parity') or {}).get('ok'), d['run_info']['setup_s'], d['run_info']['assembly']))
        print('   ', {k:v.get('ok') for k,v in d['parity']['checks'].items()})
if not ok: print(f, 'NO JSON', open(f).read()[-3000:])
PY
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=700 -k "planes-1-0 or general" > gpurun_out/pytest_multi_final.log 2>&1; echo "multi exit $?"; tail -n 3 gpurun_out/pytest_multi_final.log
