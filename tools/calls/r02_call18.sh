#!/bin/bash
# streamed path tracer (tuning key 19): parity first, then A/B against the wavefront form
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_render.py -x -q -k "streamed" 2>&1 | tail -8
for t in "" "--tuning path_stream=1"; do
  echo "== $t"
  timeout -s KILL 200 python tools/render_bench.py --spp 4 --reps 5 --no-api --same-seed $t 2>&1 | tail -1 | cut -c1-420
  timeout -s KILL 200 python tools/render_bench.py --spp 1 --reps 5 --no-api --same-seed $t 2>&1 | tail -1 | cut -c1-420
  timeout -s KILL 200 python tools/render_bench.py --width 3840 --height 2160 --spp 16 --reps 3 --no-api --same-seed $t 2>&1 | tail -1 | cut -c1-420
  timeout -s KILL 200 python tools/render_bench.py --width 3840 --height 2160 --spp 2 --reps 3 --no-api --same-seed $t 2>&1 | tail -1 | cut -c1-420
  timeout -s KILL 200 python tools/render_bench.py --spp 64 --depth 8 --reps 3 --no-api --same-seed $t 2>&1 | tail -1 | cut -c1-420
done
