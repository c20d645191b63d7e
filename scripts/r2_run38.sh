#!/bin/bash
# last check of the round: the driver's scaling command at N = 8 on the final code (slot-table gather active per rank)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $T --nproc-per-node 8 --master-port 29681 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_c5_n8.log 2>&1; echo "c5 n8 exit $?"
python - <<'PY'
import json
f='gpurun_out/bench_c5_n8.log'; ok=False
for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); ok=True; e=d.get('e2e') or {}
        print('N=%d value %.4g ms/step %.2f its %s e2e %s parity %s setup %.1fs'%(d['n_gpus'], d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], e.get('value'), (d.get('parity') or {}).get('ok'), d['run_info']['setup_s']))
        print('   ', {k:v.get('ok') for k,v in d['parity']['checks'].items()})
if not ok: print(f, 'NO JSON', open(f).read()[-3000:])
PY
