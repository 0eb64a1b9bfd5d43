#!/bin/bash
# GPU call 14 (4 GPUs): the bench line at N=4, then the reference arm under torchrun (rank 0 alone runs it)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02c14_bench_n4.json 2> gpurun_out/r02c14_bench_n4.err; echo "bench N=4 rc=$?"; grep "Error\|error" gpurun_out/r02c14_bench_n4.err | tail -5
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c14_bench_n4.json"))
for k in ("value", "ms_per_step", "e2e", "e2e_one_process", "frame_reduce", "device_render", "c3", "c4", "c5"):
    v = d.get(k)
    if isinstance(v, dict):
        v = {a: (b if not isinstance(b, (str, dict)) or len(str(b)) < 80 else str(b)[:80] + "...") for a, b in v.items()}
    print(k, json.dumps(v)[:700])
PY
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 4 --steps 5 --warmup 1 > gpurun_out/r02c14_bench_ref_n4.json 2> gpurun_out/r02c14_bench_ref_n4.err; echo "ref N=4 rc=$?"; cut -c1-400 gpurun_out/r02c14_bench_ref_n4.json
