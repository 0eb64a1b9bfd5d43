#!/bin/bash
# GPU call 10: FMNMX3 in the slab tests -- parity (full GPU suite) and the bench batch sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c10_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02c10_pytest_gpu.log
timeout 300 python tools/sweep.py "variant=3;variant=4;variant=3;variant=4" > gpurun_out/r02c10_sweep.log 2>&1; grep -E "^\{" gpurun_out/r02c10_sweep.log
timeout 300 python tools/prof_c5.py > gpurun_out/r02c10_c5_times.txt 2>&1; cat gpurun_out/r02c10_c5_times.txt
