#!/bin/bash
# Round 2, first GPU call: the questions round 1 left for the hardware (VERDICT "next round" items 4 and 9).
# Outputs -> gpurun_out/r02c1_*. Run as: gpurun --timeout 1200 -- 'bash tools/calls/r02_call1.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02c1_gpu.txt 2>&1
# 1. does the texture path / the constant path have a wavefront budget of its own? (tools/micro/l1_wavefronts.cu modes 8-15)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/l1wf.bin tools/micro/l1_wavefronts.cu && {
  ./tools/micro/l1wf.bin > gpurun_out/r02c1_l1wf_times.txt 2>&1; cat gpurun_out/r02c1_l1wf_times.txt
  timeout 400 ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_tex_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum \
    --clock-control none --csv --log-file gpurun_out/r02c1_l1wf_ncu.csv ./tools/micro/l1wf.bin > /dev/null 2>&1
}
# 2. occupancy vs spills for the default kernel: 256x5 (48 regs) against 256x6 / 128x12 (40 regs, 70 B spilled), 256x4 (54 regs)
timeout 300 python tools/sweep.py "block=256,ctas_per_sm=5;block=256,ctas_per_sm=6;block=256,ctas_per_sm=4;block=128,ctas_per_sm=12;block=128,ctas_per_sm=10;block=512,ctas_per_sm=3;block=256,ctas_per_sm=5,fetch_threshold=8;block=256,ctas_per_sm=5,fetch_threshold=24;block=256,ctas_per_sm=6,fetch_threshold=8" > gpurun_out/r02c1_sweep.log 2>&1
tail -25 gpurun_out/r02c1_sweep.log
# 3. the Whitted renderer's two queued options, once (measure, then leave it)
RACC_B200_TEST_UNMEASURED=1 timeout 300 python -m pytest tests/test_gpu_zz_whitted.py -m gpu -x -q > gpurun_out/r02c1_pytest_whitted.log 2>&1
echo "whitted pytest rc=$?" | tee -a gpurun_out/r02c1_pytest_whitted.log; tail -3 gpurun_out/r02c1_pytest_whitted.log
for t in "" "whitted_arena=1" "whitted_combine=1" "whitted_arena=1,whitted_combine=1"; do
  timeout 200 python tools/render_bench.py --whitted --depth 8 --spp 4 --reps 6 --no-api --tuning "$t" 2>&1 | tail -1 | cut -c1-900 >> gpurun_out/r02c1_whitted_options.jsonl
done
cut -c1-400 gpurun_out/r02c1_whitted_options.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c1_launches_whitted.csv \
    python tools/render_bench.py --whitted --depth 8 --spp 4 --reps 1 --no-api --tuning "whitted_arena=1,whitted_combine=1" > /dev/null 2>&1
# 4. tapered tail of the HOST staging pipeline (tuning key 17)
for t in 0 64 256; do
  RACC_B200_HOST_TAPER=$t timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({'host_taper_k': $t, 'e2e': d['e2e']['value'], 'value': d['value'], 'ok': d['e2e']['results_match_device_run']}))" >> gpurun_out/r02c1_host_taper.jsonl
done
cat gpurun_out/r02c1_host_taper.jsonl
