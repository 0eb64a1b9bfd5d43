#!/bin/bash
# GPU call 15: does a larger path batch help the deep-bounce config (C3: 64 spp, 8 bounces)? batch = 8 / 16 (default ~15) / 32 / 64 spp; lanes 2 / 4
mkdir -p gpurun_out
for b in 8 16 32 64; do
  timeout 300 python tools/render_bench.py --spp 64 --depth 8 --batch $b --reps 3 --no-api 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({'batch_spp': d['batch_spp'], 'ms_best': d['ms_best'], 'mrays_best': d['mrays_best']}))"
done
for l in 1 4; do
  RACC_B200_PATH_LANES=$l timeout 300 python tools/render_bench.py --spp 64 --depth 8 --batch 32 --reps 3 --no-api 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({'lanes': $l, 'batch_spp': d['batch_spp'], 'ms_best': d['ms_best'], 'mrays_best': d['mrays_best']}))"
done
for b in 4 16; do
  timeout 300 python tools/render_bench.py --width 3840 --height 2160 --spp 16 --depth 3 --batch $b --reps 3 --no-api 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(json.dumps({'c4 batch_spp': d['batch_spp'], 'ms_best': d['ms_best'], 'mrays_best': d['mrays_best']}))"
done
