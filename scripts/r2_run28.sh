#!/bin/bash
# the driver's scaling command at N = $1 on the final code
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node $N --master-port 2965$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_c5_n$N.log 2>&1; echo "c5 n$N exit $?"
python - $N <<'PY'
import json,sys
f='gpurun_out/bench_c5_n%s.log'%sys.argv[1]; ok=False
for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); ok=True; e=d.get('e2e') or {}
        print('N=%d value %.4g ms/step %.2f its %s e2e %s parity %s spmv %.4f ms'%(d['n_gpus'], d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], e.get('value'), (d.get('parity') or {}).get('ok'), d['roofline']['avg_launch_ms']))
if not ok: print(f, 'NO JSON', open(f).read()[-3000:])
PY
