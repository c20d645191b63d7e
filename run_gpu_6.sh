#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
TB_SPMV_COMPRESS=0 timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu_nocc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_nocc.log
for c in 1 0; do
TB_SPMV_COMPRESS=$c timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_c5_cc$c.log 2>&1
TB_SPMV_COMPRESS=$c timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_c2_cc$c.log 2>&1
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_check.py > gpurun_out/dist_check2.log 2>&1; echo "exit $?" >> gpurun_out/dist_check2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_spmv_tma -s 30 -c 2 -o gpurun_out/prof_spmv_cc python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_spmv_cc.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log; tail -n 3 gpurun_out/pytest_gpu_nocc.log; grep -E "rank [0-9]" gpurun_out/dist_check2.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*_cc*.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(f, 'value %.4g ms/step %.2f iters %.1f spmv_ms %.3f frac %.3f stored_gbs %.0f share %.3f step_frac %.3f'%(d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],r['avg_launch_ms'],r['frac'],r['stored_achieved_gbs'],r['spmv_share_of_step'],r['step_frac_of_peak']), r['bytes_per_launch']/1e9, r['stored_bytes_per_launch']/1e9)
            break
    else: print(f,'NO JSON', open(f).read()[-600:])
PY
