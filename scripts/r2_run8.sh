#!/bin/bash
# 8-GPU box: C5 at N = 8 with parity + e2e timeline + peer-wait statistics, reference arm at N = 8, C4 (general mesh) at N = 8,
# three world-4 multi-GPU test cases
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
TB_RUNHOST_TRACE=gpurun_out/runhost_trace_n8.csv timeout 900 $T --nproc-per-node 8 --master-port 29681 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_c5_n8.log 2>&1; echo "c5 n8 exit $?"
timeout 600 $T --nproc-per-node 8 --master-port 29685 bench.py --impl reference --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_ref_n8.log 2>&1; echo "ref n8 exit $?"
timeout 600 $T --nproc-per-node 8 --master-port 29683 bench.py --gpus 8 --steps 3 --warmup 3 --workload c4 --precond jacobi --e2e-steps 0 > gpurun_out/bench_c4_n8.log 2>&1; echo "c4 n8 exit $?"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=700 -k "(1-rows-1-1-4 or 1-planes-1-0-4 or general) and 4]" > gpurun_out/pytest_multi4.log 2>&1; echo "multi4 exit $?"; tail -n 5 gpurun_out/pytest_multi4.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*_n8.log'))+['gpurun_out/bench_ref_n8.log']:
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True
            e=d.get('e2e') or {}
            print(f, 'N=%d value %.4g ms/step %.2f its %s e2e %s parity %s'%(d['n_gpus'], d['value'], d['ms_per_step'], (d.get('run_info') or {}).get('cg_iters_per_step_mean'), e.get('value'), (d.get('parity') or {}).get('ok')))
            print('   ', {k:v for k,v in e.items() if k not in ('api',)})
            print('    comm', d.get('comm'))
            if d.get('parity'): print('   ', {k:(v.get('ok'), v.get('phi_rel_linf_max', v.get('max_abs_err', v.get('rel_drift', v.get('residual', v.get('max_abs_diff')))))) for k,v in d['parity']['checks'].items()})
            if d.get('cpu_baseline'): print('    cpu', d['cpu_baseline'].get('cores'), d['cpu_baseline'].get('value'))
    if not ok: print(f, 'NO JSON', open(f).read()[-1500:])
PY
