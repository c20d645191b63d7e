#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multidomain.py tests/test_ecg_leadfield.py tests/test_gpu_monodomain.py -m gpu -q --timeout=900 -k "multidomain or pacemaker or blocked or c2_full_size_1000 or poisson" > gpurun_out/pytest_md.log 2>&1; echo "md exit $?"; tail -n 30 gpurun_out/pytest_md.log
timeout 600 python scripts/bench_assembly.py > gpurun_out/bench_assembly.log 2>&1; echo "asm exit $?"; tail -n 16 gpurun_out/bench_assembly.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:k_cell_step -s 60 -c 2 -o gpurun_out/prof_cell_pcg_adaptive python bench.py --workload c2 --cell-substeps 10 --steps 2 --warmup 60 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/ncu_cell.log 2>&1; echo "ncu cell exit $?"
timeout 600 $NCU -k regex:k_bj_apply -s 20 -c 2 -o gpurun_out/prof_bj_apply python bench.py --workload c4 --precond block_jacobi --bj-rows 64 --steps 1 --warmup 1 --no-cpu --no-parity --e2e-steps 0 > gpurun_out/ncu_bj.log 2>&1; echo "ncu bj exit $?"
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -x --deselect tests/test_gpu_c5_shape.py --deselect tests/test_multidomain.py --deselect tests/test_ecg_leadfield.py -k "not c2_full_size_1000" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
