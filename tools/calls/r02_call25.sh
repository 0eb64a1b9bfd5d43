#!/bin/bash
# after the signed-zero fix in checker + kernel source: randomised parity campaign (3 seeds), ncu JSONs for the new source hash,
# every GPU test, smoke, bench + reference arm
mkdir -p gpurun_out
for seed in 1 2 3; do
  timeout -s KILL 400 python tests/fuzz/fuzz_gpu.py --seconds 150 --seed $seed > gpurun_out/r02_fuzz_gpu_seed$seed.log 2>&1; echo "fuzz seed $seed rc=$?"; tail -4 gpurun_out/r02_fuzz_gpu_seed$seed.log | cut -c1-600
done
bash tools/profile.sh > gpurun_out/r02c25_profile.log 2>&1; tail -2 gpurun_out/r02c25_profile.log
bash tools/r02_final2.sh
