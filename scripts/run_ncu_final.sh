#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_element_matrices|k_gather_rows" -s 2 -c 4 -o gpurun_out/prof_asm_gather3 python scripts/bench_assembly.py --reps 1 --cells hex --modes 2 > gpurun_out/ncu_asm_gather3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cg_persistent_tma" -s 2 -c 1 -o gpurun_out/prof_cg_persistent_tma python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu --e2e-steps 0 > gpurun_out/ncu_cgpt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_ecg_plonsey" -c 1 -o gpurun_out/prof_ecg python -m pytest tests/test_ecg.py -m gpu -q -k large > gpurun_out/ncu_ecg.log 2>&1
ls -la gpurun_out/prof_asm_gather3.ncu-rep gpurun_out/prof_cg_persistent_tma.ncu-rep gpurun_out/prof_ecg.ncu-rep
