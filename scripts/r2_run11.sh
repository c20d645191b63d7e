#!/bin/bash
# 2 GPUs: bench at N = 2 with the local structured generator, comm statistics, e2e timeline; the same with the round-1 partition path
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
TB_RUNHOST_TRACE=gpurun_out/runhost_trace_n2.csv timeout 900 $T --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c5_n2.log 2>&1; echo "c5 n2 exit $?"
TB_BENCH_LOCAL_GRID=0 timeout 900 $T --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 0 > gpurun_out/bench_c5_n2_globalgrid.log 2>&1; echo "c5 n2 global exit $?"
timeout 600 $T --master-port 29613 bench.py --gpus 2 --workload c2 --steps 20 --warmup 3 --e2e-steps 0 > gpurun_out/bench_c2_n2.log 2>&1; echo "c2 n2 exit $?"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=700 -k "2] and (general or 1-rows-1-1)" > gpurun_out/pytest_multi2b.log 2>&1; echo "multi exit $?"; tail -n 4 gpurun_out/pytest_multi2b.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*_n2*.log')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True
            e=d.get('e2e') or {}
            print(f, 'N=%d value %.4g ms/step %.2f its %s e2e %s parity %s setup %.1fs'%(d['n_gpus'], d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], e.get('value'), (d.get('parity') or {}).get('ok'), d['run_info']['setup_s']))
            print('   ', {k:v for k,v in e.items() if k not in ('api',)})
            print('    comm', d.get('comm'))
            if d.get('parity'): print('   ', {k:(v.get('ok'), v.get('phi_rel_linf_max', v.get('max_abs_err', v.get('rel_drift', v.get('residual', v.get('max_abs_diff')))))) for k,v in d['parity']['checks'].items()})
    if not ok: print(f, 'NO JSON', open(f).read()[-2500:])
PY
