#!/bin/bash
# 8 GPUs: the driver's scaling command at N = 8 and N = 4 (local structured generator, parity block, comm statistics, e2e with
# run_host trace), C4 on 8 GPUs with block-Jacobi, and the 4-rank distributed checks
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
TB_RUNHOST_TRACE=gpurun_out/runhost_trace_n8.csv timeout 900 $T --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_c5_n8.log 2>&1; echo "c5 n8 exit $?"
timeout 900 $T --nproc-per-node 4 --master-port 29622 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_c5_n4.log 2>&1; echo "c5 n4 exit $?"
timeout 600 $T --nproc-per-node 8 --master-port 29623 bench.py --gpus 8 --workload c4 --precond block_jacobi --bj-rows 64 --steps 3 --warmup 3 --e2e-steps 0 > gpurun_out/bench_c4_n8.log 2>&1; echo "c4 n8 exit $?"
timeout 600 $T --nproc-per-node 8 --master-port 29624 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_n8.log 2>&1; echo "ref n8 exit $?"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=700 -k "4]" > gpurun_out/pytest_multi4.log 2>&1; echo "multi4 exit $?"; tail -n 4 gpurun_out/pytest_multi4.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*_n[48].log'))+['gpurun_out/bench_ref_n8.log']:
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True
            e=d.get('e2e') or {}
            print(f, 'N=%d value %.4g ms/step %.2f its %s e2e %s parity %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d['run_info'].get('cg_iters_per_step_mean'), e.get('value'), (d.get('parity') or {}).get('ok')))
            print('   ', {k:v for k,v in e.items() if k not in ('api',)})
            print('    comm', {k:v for k,v in (d.get('comm') or {}).items() if k!='what'})
            print('    cpu', d.get('cpu_baseline'))
    if not ok: print(f, 'NO JSON', open(f).read()[-2500:])
PY
