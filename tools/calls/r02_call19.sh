#!/bin/bash
# streamed path tracer: refill-threshold sweep, then one full ncu capture of its kernel
mkdir -p gpurun_out
for th in 8 16 24 28 32; do
  timeout -s KILL 200 python tools/render_bench.py --spp 4 --reps 3 --no-api --same-seed --tuning path_stream=1,fetch_threshold=$th 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tuning'], d['ms_best'], d['mrays_best'])"
done
EXTRA="l1tex__data_pipe_lsu_wavefronts.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"
timeout -s KILL 600 ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:pathStreamKernel -s 2 -c 1 -o gpurun_out/prof_stream -f python tools/render_bench.py --spp 4 --reps 1 --no-api --same-seed --tuning path_stream=1 > gpurun_out/prof_stream_run.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_stream.ncu-rep > gpurun_out/prof_stream_summary.txt 2>&1
tail -3 gpurun_out/prof_stream_run.log | cut -c1-300
ls -la gpurun_out | tail -5
