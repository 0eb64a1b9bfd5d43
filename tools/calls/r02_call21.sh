#!/bin/bash
# wavefront renderer: lanes per batch (RACC_B200_PATH_LANES) on the small frames
mkdir -p gpurun_out; rm -f gpurun_out/render_bench.jsonl
for lanes in 1 2 3 4; do
  export RACC_B200_PATH_LANES=$lanes
  for cfg in "--spp 4" "--spp 1" "--width 3840 --height 2160 --spp 2" "--width 3840 --height 2160 --spp 16"; do
    timeout -s KILL 200 python tools/render_bench.py $cfg --reps 5 --no-api --same-seed 2>&1 | tail -1 | python -c "import sys,json,os; d=json.loads(sys.stdin.read()); print('lanes', os.environ['RACC_B200_PATH_LANES'], d['width'], d['spp'], d['ms_best'], d['mrays_best'])"
  done
done
