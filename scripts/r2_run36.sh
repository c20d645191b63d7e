#!/bin/bash
# ncu of the two persistent CG kernels on their configurations (C1: k_cg_persistent2, C2: k_cg_persistent_tma), final code
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:k_cg_persistent2 -s 20 -c 2 -o gpurun_out/prof_cg_persistent2_c1 python bench.py --workload c1 --steps 5 --warmup 30 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/ncu_pcg2.log 2>&1; echo "ncu c1 exit $?"
timeout 600 $NCU -k regex:k_cg_persistent_tma -s 10 -c 2 -o gpurun_out/prof_cg_persistent_tma_c2 python bench.py --workload c2 --steps 3 --warmup 12 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/ncu_pcgtma.log 2>&1; echo "ncu c2 exit $?"
