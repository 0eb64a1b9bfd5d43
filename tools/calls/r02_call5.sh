#!/bin/bash
# GPU call 5: quantised nodes with the PRMT conversion (15-bit grid): parity tests, battlefield sweep, config 5 incl. 6 CTAs per SM, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "quantised or deep_stack" > gpurun_out/r02c5_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02c5_pytest.log
timeout 300 python tools/sweep.py "variant=3;variant=4;variant=4,ctas_per_sm=6" > gpurun_out/r02c5_sweep.log 2>&1; grep -E "^\{" gpurun_out/r02c5_sweep.log
timeout 300 python tools/prof_c5.py > gpurun_out/r02c5_c5_times.txt 2>&1; cat gpurun_out/r02c5_c5_times.txt
RACC_B200_CTAS_PER_SM=6 timeout 300 python tools/prof_c5.py > gpurun_out/r02c5_c5_times_6ctas.txt 2>&1; cat gpurun_out/r02c5_c5_times_6ctas.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tracePackedKernel -s 2 -c 1 -o gpurun_out/r02c5_prof_c5_v4 -f python tools/prof_c5.py > gpurun_out/r02c5_prof_c5_v4.log 2>&1
python tools/ncu_summary.py gpurun_out/r02c5_prof_c5_v4.ncu-rep > gpurun_out/r02c5_prof_c5_v4.txt 2>&1; grep -E "time_duration|inst_executed.sum|issue_active|l1tex__throughput|pipe_alu.avg|eligible|hit_rate" gpurun_out/r02c5_prof_c5_v4.txt
