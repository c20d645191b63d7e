#!/bin/bash
# 1 GPU: round-2 ncu evidence for the default bench -- launch list (shares) and one --set full capture of the SpMV
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --no-parity > gpurun_out/ncu_launch_bench.log 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -f -k regex:k_cg_spmv_tma -s 30 -c 2 -o gpurun_out/prof_spmv_c5 python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 0 --no-parity > gpurun_out/ncu_spmv.log 2>&1; echo "ncu spmv exit $?"
timeout 600 python bench.py --workload c2 --steps 50 --warmup 5 > gpurun_out/bench_c2.log 2>&1; echo "c2 exit $?"
timeout 600 python bench.py --workload c1 --steps 200 --warmup 5 > gpurun_out/bench_c1.log 2>&1; echo "c1 exit $?"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu --precond block_jacobi --bj-rows 64 > gpurun_out/bench_c4_bj64.log 2>&1; echo "c4 exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_c2.log','gpurun_out/bench_c1.log','gpurun_out/bench_c4_bj64.log'):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); e=d.get('e2e') or {}
            print(f, 'value %.4g ms/step %.3f its %.1f e2e %s parity %s frac %s'%(d['value'], d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], e.get('value'), (d.get('parity') or {}).get('ok'), d['roofline'].get('frac')))
PY
