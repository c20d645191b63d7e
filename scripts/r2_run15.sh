#!/bin/bash
# split element kernel with coordinate prefetch: bitwise tests, in-situ launch list, bench with warm-up; traced source programs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_source_program.py tests/test_gpu_assembly.py tests/test_gpu_assembly_gather.py tests/test_gpu_c5_shape.py tests/test_lv_config4.py tests/test_ecg_leadfield.py tests/test_multidomain.py -m gpu -q --timeout=900 > gpurun_out/pytest_asm_split2.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_asm_split2.log
for v in "1 1" "0 0"; do set -- $v
TB_ELEMENT_SPLIT=$1 TB_EA_PLANES=$2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'k_element_matrices|k_gather_rows' --csv --log-file gpurun_out/launches_asm_s$1p$2.csv python scripts/bench_assembly.py --modes 2 --reps 3 --warm-s 0 > /dev/null 2>&1
python - $1 $2 <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(f'gpurun_out/launches_asm_s{sys.argv[1]}p{sys.argv[2]}.csv')) if len(r)>5 and r[0].isdigit()]
print('launch list split=%s planes=%s:'%(sys.argv[1],sys.argv[2]), ' '.join('%s=%.2f'%(r[4][5:30], float(r[-1])/ (1e6 if float(r[-1])>1e4 else 1)) for r in rows[2::8]))
PY
done
TB_ELEMENT_SPLIT=0 timeout 600 python scripts/bench_assembly.py --modes 2 > gpurun_out/bench_assembly_split0.log 2>&1; echo "asm split0 exit $?"
TB_ELEMENT_SPLIT=1 timeout 600 python scripts/bench_assembly.py --modes 2 > gpurun_out/bench_assembly_split1.log 2>&1; echo "asm split1 exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_assembly_split[01].log')):
    print(f)
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print('   %-22s %-18s %7.3f ms  %.3g el/s  frac %.3f'%(d['workload'], d['form'], d['ms'], d['elements_per_s'], d['frac']))
PY
cat gpurun_out/c4_mid_iteration_diff_*.json
