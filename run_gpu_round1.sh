#!/bin/bash
# first GPU contact: tests, smoke, bench (c2 then c5)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --workload c2 --steps 20 --warmup 3 > gpurun_out/bench_c2.log 2>&1; echo "exit $?" >> gpurun_out/bench_c2.log
timeout 600 python bench.py --workload c1 --steps 50 --warmup 3 > gpurun_out/bench_c1.log 2>&1; echo "exit $?" >> gpurun_out/bench_c1.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c5.log 2>&1; echo "exit $?" >> gpurun_out/bench_c5.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; tail -2 gpurun_out/bench_c2.log; tail -2 gpurun_out/bench_c5.log
