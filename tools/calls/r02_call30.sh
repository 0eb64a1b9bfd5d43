#!/bin/bash
# the randomised campaign with a two-device set in ONE process: scenes replicated peer to peer, HOST streams dealt over both GPUs
mkdir -p gpurun_out
RACC_B200_HOST_CHUNK=2048 timeout -s KILL 400 python tests/fuzz/fuzz_gpu.py --seconds 120 --seed 11 --devices 2 > gpurun_out/r02_fuzz_two_devices.log 2>&1; echo "fuzz rc=$?"; grep -v "^RayAccelerator" gpurun_out/r02_fuzz_two_devices.log | tail -6 | cut -c1-700
