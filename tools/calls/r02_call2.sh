#!/bin/bash
# Round 2, GPU call 2: the refactored library (per-device state, wave sizes on the device, new counters) on hardware:
# every GPU test incl. the new at-scale parity cases, smoke, the new bench line, the reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02c2_pytest_gpu.log
tail -25 gpurun_out/r02c2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c2_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r02c2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c2_bench.json 2> gpurun_out/r02c2_bench.err; echo "bench rc=$?"; cat gpurun_out/r02c2_bench.json; tail -5 gpurun_out/r02c2_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02c2_bench_ref.json 2> gpurun_out/r02c2_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r02c2_bench_ref.json; tail -3 gpurun_out/r02c2_bench_ref.err
RACC_B200_PATH_SYNC=1 timeout 200 python tools/render_bench.py --spp 16 --no-api 2>&1 | tail -1 | cut -c1-600
timeout 200 python tools/render_bench.py --spp 16 --no-api 2>&1 | tail -1 | cut -c1-600
timeout 200 python tools/render_bench.py --spp 1 --no-api 2>&1 | tail -1 | cut -c1-600
