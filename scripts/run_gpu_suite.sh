#!/bin/bash
# 1-GPU box: full -m gpu suite, smoke, bench on C5/C2/C1, launch list of the default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c5.log 2>&1
timeout 300 python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/bench_c2.log 2>&1
timeout 300 python bench.py --workload c1 --steps 50 --warmup 3 > gpurun_out/bench_c1.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launch_bench.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c[125].log'))+['gpurun_out/bench_ref.log']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            if d.get('impl')=='reference':
                print(f, 'REF value %.4g'%d['value'], d['cpu_baseline']['cores'], 'cores'); break
            r=d['roofline']
            print(f, 'value %.4g ms/step %.3f iters %.1f spmv_ms %.3f frac %.3f stored_gbs %.0f share %.3f step_frac %.3f setup %.2f'%(d['value'],d['ms_per_step'],d['config']['cg']['iters_per_step_mean'],r['avg_launch_ms'],r['frac'],r['stored_achieved_gbs'],r['spmv_share_of_step'],r['step_frac_of_peak'],d['config']['setup_s']), 'e2e %.4g cpu %.4g'%(d['e2e']['value'], d['cpu_baseline']['value']), d['clocks'])
            break
    else: print(f,'NO JSON', open(f).read()[-800:])
PY
