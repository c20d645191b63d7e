#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_mesh_pattern.py tests/test_multidomain.py tests/test_lv_config4.py tests/test_ecg_leadfield.py -m gpu -q --timeout=1200 > gpurun_out/pytest_new2.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_new2.log
