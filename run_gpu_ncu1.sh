#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
# launch list of the bench command (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c5.csv python bench.py --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_launch_bench.log 2>&1
# full capture of the dominant kernel and the two vector kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_spmv_dot -s 30 -c 2 -o gpurun_out/prof_spmv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_spmv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_cg_xr|k_cg_p|k_cell_step|k_cg_init_Mphi" -s 8 -c 4 -o gpurun_out/prof_vec python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_vec.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_assemble" -c 2 -o gpurun_out/prof_asm python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 --grid 256,256,128 > gpurun_out/ncu_asm.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; ls -la gpurun_out
