#!/bin/bash
# the campaign with large scenes in every third round (device builder's cooperative paths, deeper trees), then two ordinary seeds
mkdir -p gpurun_out
timeout -s KILL 500 python tests/fuzz/fuzz_gpu.py --seconds 240 --seed 12 --big 0.35 > gpurun_out/r02_fuzz_gpu_big.log 2>&1; echo "fuzz big rc=$?"; grep -v "^RayAccelerator" gpurun_out/r02_fuzz_gpu_big.log | tail -5 | cut -c1-700
for seed in 13 14; do
  timeout -s KILL 500 python tests/fuzz/fuzz_gpu.py --seconds 150 --seed $seed > gpurun_out/r02_fuzz_gpu_seed$seed.log 2>&1; echo "fuzz seed $seed rc=$?"; grep -v "^RayAccelerator" gpurun_out/r02_fuzz_gpu_seed$seed.log | tail -4 | cut -c1-500
done
