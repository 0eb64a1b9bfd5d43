#!/bin/bash
# the randomised campaign under compute-sanitizer initcheck (reads of uninitialised device memory) and synccheck
mkdir -p gpurun_out
timeout -s KILL 400 compute-sanitizer --tool initcheck --print-limit 20 python tests/fuzz/fuzz_gpu.py --seconds 120 --seed 15 > gpurun_out/r02_fuzz_initcheck.log 2>&1; echo "initcheck rc=$?"; grep -E "^fuzz|ERROR SUMMARY|Uninitialized|FAIL" gpurun_out/r02_fuzz_initcheck.log | head -8 | cut -c1-300
timeout -s KILL 300 compute-sanitizer --tool synccheck --print-limit 20 python tests/fuzz/fuzz_gpu.py --seconds 60 --seed 16 > gpurun_out/r02_fuzz_synccheck.log 2>&1; echo "synccheck rc=$?"; grep -E "^fuzz|ERROR SUMMARY|FAIL" gpurun_out/r02_fuzz_synccheck.log | head -6 | cut -c1-300
