#!/bin/bash
# GPU call 12: engine vs the reference's CPU query path (golden + live, 1 M rays), compute-sanitizer over the round-2 paths
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "reference_cpu_path" > gpurun_out/r02c12_pytest_cpu_path.log 2>&1; echo "pytest rc=$?"; grep -E "engine vs|passed|failed" gpurun_out/r02c12_pytest_cpu_path.log
bash tools/sanitize.sh 2>&1 | tail -12
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_render.py > gpurun_out/sanitize_render_memcheck.log 2>&1; echo "render memcheck rc=$?"; grep -E "ERROR SUMMARY|ok" gpurun_out/sanitize_render_memcheck.log | tail -3
