#!/bin/bash
# GPU call 16 (2 GPUs): multi-GPU test file incl. the rank reduce + gather through the library
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02c16_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02c16_pytest_multi.log
