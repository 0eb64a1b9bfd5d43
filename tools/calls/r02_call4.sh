#!/bin/bash
# GPU call 4: what limits the quantised-node kernel on config 5 (ncu of variant 4 alone), and its stall reasons on battlefield
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tracePackedKernel -s 2 -c 1 -o gpurun_out/r02c4_prof_c5_v4 -f python tools/prof_c5.py > gpurun_out/r02c4_prof_c5_v4.log 2>&1
python tools/ncu_summary.py gpurun_out/r02c4_prof_c5_v4.ncu-rep > gpurun_out/r02c4_prof_c5_v4.txt 2>&1; cat gpurun_out/r02c4_prof_c5_v4.txt | cut -c1-160
tail -3 gpurun_out/r02c4_prof_c5_v4.log
