#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/render_bench.py --spp 4 --reps 5 --no-api 2>&1 | tail -1 | cut -c1-400
timeout 200 python tools/render_bench.py --spp 1 --reps 5 --no-api 2>&1 | tail -1 | cut -c1-400
timeout 200 python tools/render_bench.py --width 3840 --height 2160 --spp 16 --reps 3 --no-api 2>&1 | tail -1 | cut -c1-400
timeout 200 python tools/render_bench.py --spp 64 --depth 8 --reps 3 --no-api 2>&1 | tail -1 | cut -c1-400
