#!/bin/bash
# 1-GPU checkpoint of round 2: full -m gpu suite, smoke, the driver's bench commands (both arms), C1/C2/C4 lines, block-Jacobi
# apply variants, ncu of the final split element kernel, launch list of the default bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c5.log 2>&1; echo "c5 exit $?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"
timeout 300 python bench.py --workload c2 --steps 100 --warmup 20 > gpurun_out/bench_c2.log 2>&1; echo "c2 exit $?"
timeout 300 python bench.py --workload c1 --steps 1000 --warmup 100 > gpurun_out/bench_c1.log 2>&1; echo "c1 exit $?"
for v in 0 1; do
TB_BJ_SYM=$v timeout 600 python bench.py --workload c4 --precond block_jacobi --bj-rows 64 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c4_bj64_sym$v.log 2>&1; echo "c4 sym$v exit $?"
done
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:'k_element_matrices|k_gather_rows' -c 8 -o gpurun_out/prof_asm_split_final python scripts/bench_assembly.py --modes 2 --cells hex --reps 1 --warm-s 0 > gpurun_out/ncu_asm_split.log 2>&1; echo "ncu asm exit $?"
timeout 600 $NCU -k regex:k_bj_apply -s 20 -c 2 -o gpurun_out/prof_bj_apply_sym python bench.py --workload c4 --precond block_jacobi --bj-rows 64 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/ncu_bj_sym.log 2>&1; echo "ncu bj exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launch_bench.log 2>&1; echo "launch list exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c[125].log'))+sorted(glob.glob('gpurun_out/bench_c4_bj64_sym?.log'))+['gpurun_out/bench_ref.log']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            if d.get('impl')=='reference':
                print(f, 'REF value %.4g'%d['value'], d['cpu_baseline']['cores'], 'cores'); break
            r=d['roofline']
            print(f, 'value %.4g ms/step %.4f iters %.1f frac %.3f stored_gbs %s parity %s'%(d['value'],d['ms_per_step'],d['run_info']['cg_iters_per_step_mean'],r['frac'],r.get('stored_achieved_gbs'),(d.get('parity') or {}).get('ok')), 'e2e %.4g'%((d.get('e2e') or {}).get('value') or 0), d['clocks'])
            break
    else: print(f,'NO JSON', open(f).read()[-800:])
PY
