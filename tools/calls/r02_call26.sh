#!/bin/bash
# fuzz with the quantised-node variant included, the new shared-edge regression test, full GPU suite
mkdir -p gpurun_out
timeout -s KILL 500 python tests/fuzz/fuzz_gpu.py --seconds 200 --seed 4 > gpurun_out/r02_fuzz_gpu_seed4.log 2>&1; echo "fuzz rc=$?"; grep -v "^RayAccelerator" gpurun_out/r02_fuzz_gpu_seed4.log | tail -8 | cut -c1-700
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_final_pytest_gpu.log
