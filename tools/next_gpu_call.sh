#!/bin/bash
# First GPU call after round 1: everything that was written after that round's GPU budget ran out, measured in one go.
# Outputs -> gpurun_out/next_*. Run as: gpurun --timeout 1500 -- 'bash tools/next_gpu_call.sh'
mkdir -p gpurun_out
# 0. the standing check first: all GPU tests, smoke, bench, reference arm (defaults are what was validated in round 1;
#    their SASS / control flow did not change since, and the CPU test build says the same -- confirm on hardware)
bash tools/gpu_check.sh
# 1. tuning keys 15 (grow-only wave buffers) and 16 (per-warp radiance sums) of the device Whitted renderer: parity first
RACC_B200_TEST_UNMEASURED=1 timeout 600 python -m pytest tests/test_gpu_zz_whitted.py -m gpu -x -q > gpurun_out/next_pytest_whitted.log 2>&1
echo "whitted pytest rc=$?" | tee -a gpurun_out/next_pytest_whitted.log; tail -5 gpurun_out/next_pytest_whitted.log
# 2. ... then what they buy: default / arena / combine / both, frames that differ (wave sizes change) and the same frame repeated
for t in "" "whitted_arena=1" "whitted_combine=1" "whitted_arena=1,whitted_combine=1"; do
  for s in "" "--same-seed"; do
    timeout 200 python tools/render_bench.py --whitted --depth 8 --spp 4 --reps 8 --no-api --tuning "$t" $s 2>&1 | tail -1 | cut -c1-900 >> gpurun_out/next_whitted_options.jsonl
  done
done
cat gpurun_out/next_whitted_options.jsonl | cut -c1-400
# 3. where the Whitted frame's time goes: launch list of one run per setting
for t in "" "whitted_arena=1,whitted_combine=1"; do
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "gpurun_out/next_launches_whitted_${t:-default}.csv" \
    python tools/render_bench.py --whitted --depth 8 --spp 4 --reps 1 --no-api --tuning "$t" > /dev/null 2>&1
done
# 4. does the texture path have a wavefront budget of its own? (tools/micro/l1_wavefronts.cu modes 8-10)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/l1wf.bin tools/micro/l1_wavefronts.cu && {
  ./tools/micro/l1wf.bin > gpurun_out/next_l1wf_times.txt 2>&1; cat gpurun_out/next_l1wf_times.txt
  timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_tex_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum \
    --clock-control none --csv --log-file gpurun_out/next_l1wf_ncu.csv ./tools/micro/l1wf.bin > /dev/null 2>&1
}
# 5. tapered tail of the HOST staging pipeline (tuning key 17): e2e of the bench batch with the floor at 0 (off) / 64 / 128 / 256 K rays;
#    every line carries results_match_device_run
for t in 0 64 128 256; do
  RACC_B200_HOST_TAPER=$t timeout 300 python bench.py --steps 20 --warmup 3 2> /dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({'host_taper_k': $t, 'e2e': d['e2e']}))" >> gpurun_out/next_host_taper.jsonl
done
cat gpurun_out/next_host_taper.jsonl | cut -c1-300
