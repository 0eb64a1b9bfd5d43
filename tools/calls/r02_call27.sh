#!/bin/bash
# after the quantised-grid fix: fuzz (two seeds, quantised variant included), ncu JSONs for the new source hash, the closing pass
mkdir -p gpurun_out
for seed in 5 6; do
  timeout -s KILL 400 python tests/fuzz/fuzz_gpu.py --seconds 150 --seed $seed > gpurun_out/r02_fuzz_gpu_seed$seed.log 2>&1; echo "fuzz seed $seed rc=$?"; grep -v "^RayAccelerator" gpurun_out/r02_fuzz_gpu_seed$seed.log | tail -6 | cut -c1-700
done
bash tools/profile.sh > gpurun_out/r02c27_profile.log 2>&1; tail -2 gpurun_out/r02c27_profile.log
bash tools/r02_final2.sh
