#!/bin/bash
mkdir -p gpurun_out
# memory checker over the smoke path (assembly gather, persistent CG, cell sweep) and the multi-kernel CG path
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.log 2>&1; echo "memcheck smoke exit $?"; tail -3 gpurun_out/sanitizer_smoke.log
TB_CG_PERSISTENT=0 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke_mk.log 2>&1; echo "memcheck smoke (multi-kernel CG) exit $?"; tail -3 gpurun_out/sanitizer_smoke_mk.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_assembly_gather.py -m gpu -q -x -k "matrices_bitwise or fallback" > gpurun_out/sanitizer_asm.log 2>&1; echo "memcheck assembly exit $?"; tail -3 gpurun_out/sanitizer_asm.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_assembly_gather.py -m gpu -q -x -k "matrices_bitwise and HEX8-nel1" > gpurun_out/sanitizer_race.log 2>&1; echo "racecheck assembly exit $?"; tail -3 gpurun_out/sanitizer_race.log
# ncu full captures of the new kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_element_matrices|k_gather_rows" -c 4 -o gpurun_out/prof_asm_gather2 python scripts/bench_assembly.py --reps 1 --cells hex --modes 2 > gpurun_out/ncu_asm_gather2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cg_persistent" -s 3 -c 2 -o gpurun_out/prof_cg_persistent python bench.py --workload c1 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/ncu_cgp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell_step" -c 2 -o gpurun_out/prof_cell_fhn python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/ncu_cell_fhn.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
