#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_lv2.py > gpurun_out/debug_lv2.log 2>&1; tail -40 gpurun_out/debug_lv2.log | cut -c1-330
timeout 300 python -m pytest tests/test_gpu_assembly_gather.py -m gpu -q -x --timeout=300 2>&1 | tail -3
timeout 600 python scripts/bench_assembly.py --cells hex > gpurun_out/bench_assembly.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/bench_assembly.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['workload'], d['form'], 'mode', d['mode_requested'], d['mode_used'], 'chunks', d['chunks'], 'ms %.2f'%d['ms'], 'Mel/s %.1f'%(d['elements_per_s']/1e6), 'GB/s %.0f frac %.3f'%(d['achieved_gbs'], d['frac']))
    else: print(l.rstrip()[-300:])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_element_matrices|k_gather_rows" -c 4 -o gpurun_out/prof_asm_gather python scripts/bench_assembly.py --reps 1 --cells hex --modes 2 > gpurun_out/ncu_asm_gather.log 2>&1
ls -la gpurun_out/prof_asm_gather.ncu-rep
