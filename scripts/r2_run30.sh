#!/bin/bash
# gather with four adjacency entries in flight per lane: bitwise tests, launch list, C3 bench (hex; tets with the slot table too)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_assembly_gather.py tests/test_gpu_c5_shape.py tests/test_ecg_leadfield.py tests/test_multidomain.py -m gpu -q --timeout=900 > gpurun_out/pytest_r2j.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_r2j.log
TB_GATHER_POS_ALL=1 timeout 900 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_assembly_gather.py -m gpu -q --timeout=900 > gpurun_out/pytest_r2j_all.log 2>&1; echo "pytest (all types) exit $?"; tail -n 2 gpurun_out/pytest_r2j_all.log
for v in 0 1; do
TB_GATHER_POS_ALL=$v timeout 600 python scripts/bench_assembly.py --modes 2 > gpurun_out/bench_assembly_gb$v.log 2>&1; echo "asm all$v exit $?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_assembly_gb[01].log')):
    print(f)
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l)
            if d['form'] in ('mass','diffusion_tensor'): print('   %-22s %-18s %7.3f ms  %.3g el/s  frac %.3f'%(d['workload'], d['form'], d['ms'], d['elements_per_s'], d['frac']))
PY
TB_GATHER_POS_ALL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'k_gather_rows' --csv --log-file gpurun_out/launches_asm_gb.csv python scripts/bench_assembly.py --modes 2 --reps 3 --warm-s 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_asm_gb.csv')) if len(r)>5 and r[0].isdigit()]
seen={}
for r in rows:
    name=r[4].replace('void ','')[:34]; t=float(r[-1]); t=t/1e6 if t>1e4 else t
    seen.setdefault(name,[]).append(t)
for k,v in seen.items(): print('   %-36s n=%d median %.3f ms'%(k,len(v),sorted(v)[len(v)//2]))
PY
