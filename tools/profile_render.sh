#!/bin/bash
# ncu evidence for the device-side renderers (run under gpurun, one GPU). Outputs land in gpurun_out/.
#   bash tools/profile_render.sh            path tracer
#   bash tools/profile_render.sh --whitted  Whitted renderer (first thing to look at next: whittedShadeKernel's atomics)
mkdir -p gpurun_out
if [ "$1" == "--whitted" ]; then ARGS="--whitted --depth 8 --spp 4"; K=whittedShadeKernel; TAG=whitted; else ARGS="--spp 4"; K=pathShadeKernel; TAG=path; fi
CMD="python tools/render_bench.py --reps 1 --no-api $ARGS"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/render_launches_$TAG.csv $CMD > gpurun_out/render_launches_$TAG.log 2>&1
# the shading kernel, full set: the first frame set's launches (wave 0 first)
ncu --set full --clock-control none --import-source on -k regex:$K -c 4 -o gpurun_out/prof_render_$TAG -f $CMD > gpurun_out/prof_render_$TAG.log 2>&1
ls -la gpurun_out | grep -E "render_launches_$TAG|prof_render_$TAG"
