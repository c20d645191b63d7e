#!/bin/bash
# 8 GPUs, final: the driver's scaling command at N = 8 (parity block, comm statistics, e2e + chunk-count sweep)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
TB_RUNHOST_TRACE=gpurun_out/runhost_trace_n8.csv timeout 900 $T --nproc-per-node 8 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 5 --e2e-chunk-sweep 2,4,8,32,64 > gpurun_out/bench_c5_n8.log 2>&1; echo "c5 n8 exit $?"
python - <<'PY'
import json
f='gpurun_out/bench_c5_n8.log'; ok=False
for l in open(f):
    if l.startswith('{'):
        d=json.loads(l); ok=True; e=d.get('e2e') or {}
        print('N=%d value %.4g ms/step %.2f its %s e2e %s parity %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], e.get('value'), (d.get('parity') or {}).get('ok')))
        print('   ', {k:v for k,v in e.items() if k not in ('api',)})
        print('    comm', {k:v for k,v in (d.get('comm') or {}).items() if k!='what'})
        print('    roofline', {k:d['roofline'][k] for k in ('avg_launch_ms','stored_achieved_gbs','frac','step_frac_of_peak')})
if not ok: print(f, 'NO JSON', open(f).read()[-3000:])
PY
