#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (engine + reference arm), API throughput. Outputs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout 300 ./oracle/_ref/racc_render_gpu --width 1920 --height 1080 --frames 8 > gpurun_out/render_path.json 2>&1; cat gpurun_out/render_path.json
timeout 300 ./oracle/_ref/racc_render_gpu --whitted --width 1920 --height 1080 --frames 8 --out gpurun_out/whitted.ppm > gpurun_out/render_whitted.json 2>&1; cat gpurun_out/render_whitted.json
# the example renderers' frames rendered on the device (csrc/pathtrace.cu, csrc/whitted.cu) beside the host-shaded API path
timeout 200 python tools/render_bench.py --spp 16 > gpurun_out/render_device_path.json 2>&1; tail -1 gpurun_out/render_device_path.json | cut -c1-700
timeout 200 python tools/render_bench.py --whitted --depth 8 --spp 4 --reps 5 > gpurun_out/render_device_whitted.json 2>&1; tail -1 gpurun_out/render_device_whitted.json | cut -c1-700
