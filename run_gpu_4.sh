#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for v in 1 2 3 4; do
TB_SPMV_VARIANT=$v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_c5_w$v.log 2>&1
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > gpurun_out/dist_check2.log 2>&1; echo "exit $?" >> gpurun_out/dist_check2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c5_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/bench_c5_2gpu.log
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 4 gpurun_out/dist_check2.log; tail -n 3 gpurun_out/bench_c5_2gpu.log | cut -c1-600
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c5_w*.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(f, 'value %.4g ms/step %.2f iters %.1f spmv_ms %.3f frac %.3f share %.3f step_frac %.3f'%(d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],r['avg_launch_ms'],r['frac'],r['spmv_share_of_step'],r['step_frac_of_peak']))
            break
    else: print(f,'NO JSON', open(f).read()[-300:])
PY
