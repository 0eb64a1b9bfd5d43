#!/bin/bash
# Round 2, closing single-GPU pass after the header split (kernel source hash changed, SASS identical): ncu JSONs again,
# every GPU test, smoke, bench + reference arm, wavefront vs streamed renderer, sanitizer over both renderer forms
mkdir -p gpurun_out
rm -f gpurun_out/render_bench.jsonl
bash tools/profile.sh > gpurun_out/r02c20_profile.log 2>&1; tail -3 gpurun_out/r02c20_profile.log
bash tools/r02_final2.sh
for t in "" "--tuning path_stream=1"; do
  timeout -s KILL 200 python tools/render_bench.py --spp 4 --reps 5 --no-api --same-seed $t > /dev/null 2>&1
  timeout -s KILL 200 python tools/render_bench.py --spp 1 --reps 5 --no-api --same-seed $t > /dev/null 2>&1
  timeout -s KILL 200 python tools/render_bench.py --width 3840 --height 2160 --spp 16 --reps 3 --no-api --same-seed $t > /dev/null 2>&1
  timeout -s KILL 200 python tools/render_bench.py --width 3840 --height 2160 --spp 2 --reps 3 --no-api --same-seed $t > /dev/null 2>&1
  timeout -s KILL 200 python tools/render_bench.py --spp 64 --depth 8 --reps 3 --no-api --same-seed $t > /dev/null 2>&1
done
python -c "
import json
for l in open('gpurun_out/render_bench.jsonl'):
    d=json.loads(l); print(d['width'],d['height'],d['spp'],d['max_depth'],d['tuning'],d['ms_best'],d['mrays_best'])
"
timeout -s KILL 900 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_render.py > gpurun_out/sanitize_render_memcheck.log 2>&1; echo "render memcheck rc=$?"; grep -E "ERROR SUMMARY|ok" gpurun_out/sanitize_render_memcheck.log | tail -4
timeout -s KILL 900 compute-sanitizer --tool racecheck --print-limit 10 python tools/sanitize_render.py > gpurun_out/sanitize_render_racecheck.log 2>&1; echo "render racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ok" gpurun_out/sanitize_render_racecheck.log | tail -4
