#!/bin/bash
# fuzz, extended: larger scenes now and then, counted launches over ragged DEVICE streams, Whitted frames
mkdir -p gpurun_out
for seed in 7 8; do
  timeout -s KILL 500 python tests/fuzz/fuzz_gpu.py --seconds 170 --seed $seed > gpurun_out/r02_fuzz_gpu_seed$seed.log 2>&1; echo "fuzz seed $seed rc=$?"; grep -v "^RayAccelerator" gpurun_out/r02_fuzz_gpu_seed$seed.log | tail -6 | cut -c1-900
done
