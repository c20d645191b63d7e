#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_lv_config4.py -m gpu -q -x --timeout=300 2>&1 | tail -5
for i in 1 2; do
timeout 600 python scripts/bench_assembly.py --cells hex --modes 2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['form'], 'mode', d['mode_used'], 'ms %.2f'%d['ms'])
    else: print(l.rstrip()[-200:])"
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_asm.csv python scripts/bench_assembly.py --reps 3 --cells hex --modes 2 > gpurun_out/ncu_asm_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_asm.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
for r in rows[1:]:
    if 'k_element' in r[ki] or 'k_gather' in r[ki]: print(r[ki][:40], r[vi], r[ui])
PY
