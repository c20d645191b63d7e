#!/bin/bash
# split element kernel (geometry per (element, point) + a thread per column pair) and plane-major EA: bitwise tests, C3 bench in
# both variants, ncu of the new kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_assembly.py tests/test_gpu_assembly_gather.py tests/test_gpu_c5_shape.py tests/test_lv_config4.py tests/test_ecg_leadfield.py -m gpu -q --timeout=900 -x > gpurun_out/pytest_asm_split.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_asm_split.log
TB_ELEMENT_SPLIT=0 timeout 600 python scripts/bench_assembly.py --modes 2 > gpurun_out/bench_assembly_split0.log 2>&1; echo "asm split0 exit $?"
TB_ELEMENT_SPLIT=1 timeout 600 python scripts/bench_assembly.py --modes 2 > gpurun_out/bench_assembly_split1.log 2>&1; echo "asm split1 exit $?"
TB_ELEMENT_SPLIT=1 TB_ELEMENT_QB=8 timeout 600 python scripts/bench_assembly.py --modes 2 --cells hex > gpurun_out/bench_assembly_split1_qb8.log 2>&1; echo "asm split1 qb8 exit $?"
TB_ELEMENT_SPLIT=1 TB_ELEMENT_QB=2 timeout 600 python scripts/bench_assembly.py --modes 2 --cells hex > gpurun_out/bench_assembly_split1_qb2.log 2>&1; echo "asm split1 qb2 exit $?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_assembly_split*.log')):
    print(f)
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print('   %-22s %-18s %7.3f ms  %.3g el/s  frac %.3f'%(d['workload'], d['form'], d['ms'], d['elements_per_s'], d['frac']))
PY
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:'k_element_matrices|k_gather_rows' -c 8 -o gpurun_out/prof_asm_split python scripts/bench_assembly.py --modes 2 --cells hex --reps 1 > gpurun_out/ncu_asm_split.log 2>&1; echo "ncu exit $?"
