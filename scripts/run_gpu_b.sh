#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_lv_config4.py -m gpu -q -x --timeout=300 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_asm.csv python scripts/bench_assembly.py --reps 1 --cells hex --modes 2 > gpurun_out/ncu_asm_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_asm.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:]:
    print(r[ki][:60], r[vi], r[ui])
PY
