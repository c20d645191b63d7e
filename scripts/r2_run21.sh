#!/bin/bash
# header-only column staging in the TMA SpMV (up to 32 warps per SM): tests, C5 / C2 / C4 with TB_SPMV_HDRONLY = 0 / 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_spmv_cg.py tests/test_gpu_c5_shape.py tests/test_gpu_monodomain.py tests/test_gpu_fusep.py tests/test_gpu_exact_dot.py tests/test_lv_config4.py -m gpu -q --timeout=900 > gpurun_out/pytest_r2e.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_r2e.log
for v in 0 1; do
TB_SPMV_HDRONLY=$v timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/bench_c5_hdr$v.log 2>&1; echo "c5 hdr$v exit $?"
TB_SPMV_HDRONLY=$v timeout 300 python bench.py --workload c2 --steps 100 --warmup 20 --no-cpu --e2e-steps 0 > gpurun_out/bench_c2_hdr$v.log 2>&1; echo "c2 hdr$v exit $?"
TB_SPMV_HDRONLY=$v timeout 300 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu --e2e-steps 0 --no-parity > gpurun_out/bench_c4_hdr$v.log 2>&1; echo "c4 hdr$v exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_c*_hdr?.log')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print(f, 'ms/step %.4f its %.1f spmv %.4f ms stored %.0f GB/s frac %.3f parity %s'%(d['ms_per_step'], d['run_info']['cg_iters_per_step_mean'], r['avg_launch_ms'], r.get('stored_achieved_gbs') or 0, r['frac'], (d.get('parity') or {}).get('ok')))
            break
    else: print(f, 'NO JSON', open(f).read()[-1500:])
PY
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 900 $NCU -k regex:k_cg_spmv_tma -s 30 -c 2 -o gpurun_out/prof_spmv_c5_hdr python bench.py --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/ncu_spmv_hdr.log 2>&1; echo "ncu exit $?"
