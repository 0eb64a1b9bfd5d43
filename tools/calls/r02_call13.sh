#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "reference_cpu_path" > gpurun_out/r02c13_pytest_cpu_path.log 2>&1; echo "pytest rc=$?"; grep -E "engine vs|passed|failed|Error" gpurun_out/r02c13_pytest_cpu_path.log
