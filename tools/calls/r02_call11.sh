#!/bin/bash
# GPU call 11: packed node copy renumbered for parent/child line sharing (tuning key 19) -- A/B on the bench batch and config 5
mkdir -p gpurun_out
timeout 300 python tools/sweep.py "variant=3;variant=3" > gpurun_out/r02c11_sweep_layout0.log 2>&1; grep -E "^\{" gpurun_out/r02c11_sweep_layout0.log
RACC_B200_NODE_LAYOUT=1 timeout 300 python tools/sweep.py "variant=3;variant=3" > gpurun_out/r02c11_sweep_layout1.log 2>&1; grep -E "^\{" gpurun_out/r02c11_sweep_layout1.log
RACC_B200_NODE_LAYOUT=1 timeout 300 python tools/prof_c5.py > gpurun_out/r02c11_c5_layout1.txt 2>&1; cat gpurun_out/r02c11_c5_layout1.txt
RACC_B200_NODE_LAYOUT=1 timeout 400 ncu --set full --clock-control none -k regex:tracePackedKernel -s 2 -c 2 -o gpurun_out/r02c11_prof_layout1 -f python tools/prof_case.py variant=3 2 > gpurun_out/r02c11_prof_layout1.log 2>&1
python tools/ncu_summary.py gpurun_out/r02c11_prof_layout1.ncu-rep | grep -E "kernel|time_duration|l1tex__t_sector_hit|lts__t_sector_hit|l1tex__throughput|issue_active|lts__throughput"
