#!/bin/bash
# GPU call 9 (8 GPUs): the bench line at N=8 -- weak scaling, engine-level frame reduction over 8 ranks, e2e against the box's
# pinned-copy ceiling with all 8 GPUs copying, e2e_one_process over 8 devices, c4 strong scaling, c5 split 8 ways.
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02c9_bench_n8.json 2> gpurun_out/r02c9_bench_n8.err; echo "bench N=8 rc=$?"; grep "bench \|Error\|error" gpurun_out/r02c9_bench_n8.err | tail -12
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c9_bench_n8.json"))
for k in ("value", "ms_per_step", "e2e", "e2e_one_process", "frame_reduce", "device_render", "c3", "c4", "c5"):
    v = d.get(k)
    if isinstance(v, dict):
        v = {a: (b if not isinstance(b, str) or len(b) < 80 else b[:80] + "...") for a, b in v.items()}
    print(k, json.dumps(v)[:1300])
PY
