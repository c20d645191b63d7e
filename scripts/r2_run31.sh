#!/bin/bash
# split element kernel with cp.async double-buffered coordinates: bitwise tests, launch list, C3 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_assembly_gather.py tests/test_gpu_c5_shape.py tests/test_ecg_leadfield.py tests/test_multidomain.py tests/test_lv_config4.py -m gpu -q --timeout=900 -k "not 1000_steps" > gpurun_out/pytest_r2k.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_r2k.log
timeout 600 python scripts/bench_assembly.py --modes 2 > gpurun_out/bench_assembly_cpasync.log 2>&1; echo "asm exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_assembly_cpasync.log'):
    if l.startswith('{'):
        d=json.loads(l)
        if d['form'] in ('mass','diffusion_tensor'): print('   %-22s %-18s %7.3f ms  %.3g el/s  frac %.3f'%(d['workload'], d['form'], d['ms'], d['elements_per_s'], d['frac']))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'k_element_matrices|k_gather_rows' --csv --log-file gpurun_out/launches_asm_cpasync.csv python scripts/bench_assembly.py --modes 2 --reps 3 --warm-s 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_asm_cpasync.csv')) if len(r)>5 and r[0].isdigit()]
seen={}
for r in rows:
    name=r[4].replace('void ','')[:40]; t=float(r[-1]); t=t/1e6 if t>1e4 else t
    seen.setdefault(name,[]).append(t)
for k,v in seen.items(): print('   %-42s n=%d median %.3f ms'%(k,len(v),sorted(v)[len(v)//2]))
PY
