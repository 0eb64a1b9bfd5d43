#!/bin/bash
# the randomised campaign under compute-sanitizer (memcheck, then racecheck): out-of-bounds or racy accesses on degenerate scenes
mkdir -p gpurun_out
timeout -s KILL 600 compute-sanitizer --tool memcheck --print-limit 20 python tests/fuzz/fuzz_gpu.py --seconds 150 --seed 9 > gpurun_out/r02_fuzz_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "^fuzz|ERROR SUMMARY|Invalid|FAIL" gpurun_out/r02_fuzz_memcheck.log | head -12 | cut -c1-400
timeout -s KILL 600 compute-sanitizer --tool racecheck --print-limit 20 python tests/fuzz/fuzz_gpu.py --seconds 120 --seed 10 > gpurun_out/r02_fuzz_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "^fuzz|RACECHECK SUMMARY|hazard|FAIL" gpurun_out/r02_fuzz_racecheck.log | head -12 | cut -c1-400
