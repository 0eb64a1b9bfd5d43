#!/bin/bash
# randomised parity campaign (tests/fuzz/fuzz_gpu.py), then the ncu JSONs for the current kernel-source hash
mkdir -p gpurun_out
timeout -s KILL 500 python tests/fuzz/fuzz_gpu.py --seconds 300 --seed 1 > gpurun_out/r02_fuzz_gpu.log 2>&1; echo "fuzz rc=$?"; tail -8 gpurun_out/r02_fuzz_gpu.log | cut -c1-400
bash tools/profile.sh > gpurun_out/r02c24_profile.log 2>&1; tail -2 gpurun_out/r02c24_profile.log
python -c "
import json
for n in ('ncu_bench_traffic','ncu_c5_traffic'):
    d=json.load(open('gpurun_out/'+n+'.json')); print(n, d['kernel_source_sha16'], d['dram_bytes_per_launch'], d['l1_wavefronts_per_launch'], d['limiters'])
"
