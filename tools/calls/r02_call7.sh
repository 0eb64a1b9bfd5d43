#!/bin/bash
# GPU call 7 (2 GPUs): several B200s behind one process (library-level multi-device, NCCL frame reduction, racc::cudaDevices),
# then the bench line at N=2 (engine communicator across ranks, e2e_one_process, c4 strong scaling, c5 split over ranks).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r02c7_gpus.txt 2>&1; cat gpurun_out/r02c7_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/r02c7_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02c7_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02c7_bench_n2.json 2> gpurun_out/r02c7_bench_n2.err; echo "bench N=2 rc=$?"; tail -5 gpurun_out/r02c7_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c7_bench_n2.json"))
for k in ("value", "ms_per_step", "e2e", "e2e_one_process", "frame_reduce", "device_render", "c3", "c4", "c5"):
    v = d.get(k)
    if isinstance(v, dict):
        v = {a: (b if not isinstance(b, str) or len(b) < 80 else b[:80] + "...") for a, b in v.items()}
    print(k, json.dumps(v)[:900])
PY
