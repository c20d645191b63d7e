#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for n in 1 2 4; do
if [ $n -eq 1 ]; then
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/scale_c5_$n.log 2>&1
else
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_c5_$n.log 2>&1
fi
echo "exit $?" >> gpurun_out/scale_c5_$n.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_spmv_tma -s 30 -c 2 -o gpurun_out/prof_spmv_tma python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_spmv.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/scale_c5_*.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(f, 'n %d value %.4g ms/step %.2f iters %.1f spmv_ms %.3f frac %.3f share %.3f e2e %s cpu %s'%(d['n_gpus'],d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],r['avg_launch_ms'],r['frac'],r['spmv_share_of_step'], d['e2e'] and d['e2e']['value'], d['cpu_baseline'] and d['cpu_baseline']['value']))
            break
    else: print(f,'NO JSON', open(f).read()[-600:])
PY
