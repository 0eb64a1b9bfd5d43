#!/bin/bash
# GPU call 8 (2 GPUs): the bench line at N=2 after the frame-reduce fix
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02c8_bench_n2.json 2> gpurun_out/r02c8_bench_n2.err; echo "bench N=2 rc=$?"; grep "bench " gpurun_out/r02c8_bench_n2.err | tail -5
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c8_bench_n2.json"))
for k in ("value", "ms_per_step", "e2e", "e2e_one_process", "frame_reduce", "device_render", "c3", "c4", "c5"):
    v = d.get(k)
    if isinstance(v, dict):
        v = {a: (b if not isinstance(b, str) or len(b) < 80 else b[:80] + "...") for a, b in v.items()}
    print(k, json.dumps(v)[:1200])
PY
