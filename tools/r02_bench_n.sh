#!/bin/bash
# The bench line at N GPUs of one box (run under `gpurun --gpus N`): bash tools/r02_bench_n.sh N [with-multi-gpu-tests]
N=${1:-2}
mkdir -p gpurun_out
if [ -n "$2" ]; then
  timeout -s KILL 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_pytest_gpu_multi.log 2>&1; echo "multi-GPU pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_multi.log
fi
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"; grep "Error\|error" gpurun_out/r02_bench_n$N.err | tail -5
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_n$N.json"))
for k in ("value", "ms_per_step", "e2e", "e2e_one_process", "frame_reduce", "device_render", "c3", "c4", "c5"):
    v = d.get(k)
    if isinstance(v, dict):
        v = {a: (b if not isinstance(b, str) or len(b) < 60 else b[:60] + "...") for a, b in v.items() if a not in ("what", "note")}
    print(k, json.dumps(v)[:700])
PY
