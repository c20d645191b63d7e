#!/bin/bash
# final 1-GPU checkpoint of round 2: full -m gpu suite, smoke, both bench arms, C1/C2/C4 lines, ncu of the final SpMV and gather
# kernels, launch list, memcheck of the kernels added after the first sanitizer run
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c5.log 2>&1; echo "c5 exit $?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1; echo "ref exit $?"
timeout 300 python bench.py --workload c2 --steps 100 --warmup 20 > gpurun_out/bench_c2.log 2>&1; echo "c2 exit $?"
timeout 300 python bench.py --workload c1 --steps 1000 --warmup 100 > gpurun_out/bench_c1.log 2>&1; echo "c1 exit $?"
timeout 600 python bench.py --workload c4 --precond block_jacobi --bj-rows 64 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c4_bj64.log 2>&1; echo "c4 exit $?"
python - <<'PY'
import json
for f in ['gpurun_out/bench_c5.log','gpurun_out/bench_c2.log','gpurun_out/bench_c1.log','gpurun_out/bench_c4_bj64.log','gpurun_out/bench_ref.log']:
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
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:k_cg_spmv_tma -s 30 -c 2 -o gpurun_out/prof_spmv_c5_final python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/ncu_spmv_final.log 2>&1; echo "ncu spmv exit $?"
timeout 600 $NCU -k regex:'k_gather_rows_pos|k_adjpos_build' -c 4 -o gpurun_out/prof_gather_pos python scripts/bench_assembly.py --modes 2 --cells hex --reps 1 --warm-s 0 > gpurun_out/ncu_gather_pos.log 2>&1; echo "ncu gather exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launch_bench.log 2>&1; echo "launch list exit $?"
S=/usr/local/cuda/bin/compute-sanitizer
{
echo "## memcheck (final code): assembly with the slot-table gather, all cell types"; echo '```'
timeout 600 $S --tool memcheck python -m pytest tests/test_gpu_assembly_gather.py tests/test_gpu_assembly.py -m gpu -q --timeout=500 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -6; echo '```'
echo "## memcheck (final code): TMA sweep with prefetched slice metadata, header-only stage, wide rows (SpMV / CG tests, LV apex rows)"; echo '```'
timeout 900 $S --tool memcheck python -m pytest tests/test_gpu_spmv_cg.py tests/test_lv_config4.py -m gpu -q --timeout=800 -k "bitwise or apex or matches_oracle" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | head -6; echo '```'
} > gpurun_out/compute_sanitizer_r02_final.md 2>&1
cat gpurun_out/compute_sanitizer_r02_final.md
