#!/bin/bash
# ncu evidence for the bench command (run under gpurun, one GPU). Outputs land in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches_run.log 2>&1
# the top kernel, full set: -s 7 skips the 4 counted roofline-accounting launches and the 3 warm-ups,
# -c 2 captures the two timed launches
ncu --set full --clock-control none --import-source on -k regex:tracePackedKernel -s 7 -c 2 -o gpurun_out/prof -f $CMD > gpurun_out/prof_run.log 2>&1
ls -la gpurun_out
