#!/bin/bash
# 2-GPU box: multi-GPU tests (world 2), C5 and C4 at N=2 with the parity block, block-Jacobi apply timing on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=700 -k "2-" > gpurun_out/pytest_multi2.log 2>&1; echo "multi exit $?"; tail -n 12 gpurun_out/pytest_multi2.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $T --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c5_n2.log 2>&1; echo "c5 n2 exit $?"
timeout 600 $T --master-port 29612 bench.py --gpus 2 --steps 3 --warmup 3 --workload c4 --precond jacobi --e2e-steps 0 > gpurun_out/bench_c4_n2.log 2>&1; echo "c4 n2 exit $?"
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --precond block_jacobi --bj-rows 64 > gpurun_out/bench_c4_bj64.log 2>&1; echo "c4 bj exit $?"
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 --precond block_jacobi --bj-rows 32 > gpurun_out/bench_c4_bj32.log 2>&1; echo "c4 bj32 exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*.log')):
    ok=False
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); ok=True
            e=d.get('e2e') or {}
            print(f, 'N=%d value %.4g ms/step %.2f its %.1f e2e %s parity %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], e.get('value'), (d.get('parity') or {}).get('ok')))
            print('   ', d['run_info']['parallelism'][:150])
            if d.get('parity'): print('   ', {k:(v.get('ok'), v.get('phi_rel_linf_max', v.get('max_abs_err', v.get('rel_drift', v.get('residual', v.get('max_abs_diff')))))) for k,v in d['parity']['checks'].items()})
    if not ok: print(f, 'NO JSON', open(f).read()[-1500:])
PY
