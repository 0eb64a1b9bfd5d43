#!/bin/bash
# comments in the kernel source changed (a path): ncu JSONs for the new source hash, then the closing single-GPU pass once more
mkdir -p gpurun_out
bash tools/profile.sh > gpurun_out/r02c32_profile.log 2>&1; tail -2 gpurun_out/r02c32_profile.log
bash tools/r02_final2.sh
