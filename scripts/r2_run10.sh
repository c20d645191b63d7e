#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_io.py tests/test_gpu_mesh_pattern.py -m gpu -q --timeout=600 > gpurun_out/pytest_io.log 2>&1; echo "io exit $?"; tail -n 5 gpurun_out/pytest_io.log
bash scripts/r2_run9.sh
