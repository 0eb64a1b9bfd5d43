#!/bin/bash
# wavefront renderer: traversal launches capped at N CTAs per SM so that the other lane's shading kernel runs beside them
mkdir -p gpurun_out; rm -f gpurun_out/render_bench.jsonl
for ctas in 0 4 3; do
  for cfg in "--spp 4" "--spp 1" "--width 3840 --height 2160 --spp 2" "--width 3840 --height 2160 --spp 16" "--spp 64 --depth 8"; do
    timeout -s KILL 200 python tools/render_bench.py $cfg --reps 4 --no-api --same-seed --tuning path_trace_ctas=$ctas 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['tuning'], d['width'], d['spp'], d['max_depth'], d['ms_best'], d['mrays_best'])"
  done
done
timeout -s KILL 600 python -m pytest tests/test_gpu_render.py -x -q 2>&1 | tail -3
