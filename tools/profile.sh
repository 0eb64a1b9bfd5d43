#!/bin/bash
# ncu evidence for the bench command (run under gpurun, one GPU). Outputs land in gpurun_out/; the two JSON files bench.py
# quotes (DRAM traffic, L1 wavefronts, limiter percentages + the hash of the kernel source they describe) are written by
# tools/ncu_traffic.py ON THE BOX, so that the hash is the snapshot's, and copied to gpurun_out/ for profiles/.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --c5-triangles 0"
EXTRA="l1tex__data_pipe_lsu_wavefronts.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches_run.log 2>&1
# the top kernel, full set: -s 7 skips the 4 counted roofline-accounting launches and the 3 warm-ups,
# -c 2 captures the two timed launches
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:tracePackedKernel -s 7 -c 2 -o gpurun_out/prof -f $CMD > gpurun_out/prof_run.log 2>&1
python tools/ncu_traffic.py gpurun_out/prof.ncu-rep ncu_bench_traffic && cp profiles/ncu_bench_traffic.json gpurun_out/
python tools/ncu_summary.py gpurun_out/prof.ncu-rep > gpurun_out/prof_summary.txt 2>&1
# config 5: one pass of the exact format (the bench's c5 record), then the quantised one
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:tracePackedKernel -s 1 -c 2 -o gpurun_out/prof_c5 -f python tools/prof_c5.py > gpurun_out/prof_c5_run.log 2>&1
python tools/ncu_traffic.py gpurun_out/prof_c5.ncu-rep ncu_c5_traffic "256, 5, 16, 0>" && cp profiles/ncu_c5_traffic.json gpurun_out/
python tools/ncu_summary.py gpurun_out/prof_c5.ncu-rep > gpurun_out/prof_c5_summary.txt 2>&1
ls -la gpurun_out
