"""BASELINE.json config 5 across GPUs: synthetic 10 M-triangle soup, 100 M uniform-random rays, 1 vs N B200
(strong scaling: the ray set is fixed and split by index; the scene is built on every GPU). Development tool.

    python tools/c5_multi.py                                              (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c5_multi.py

Device time of the traversal launch (re-binning included) by CUDA events, max over ranks; algorithmic bytes from
the kernel's visit counters; one NCCL all-reduce of the frame counters. Appends a JSON line to gpurun_out/c5_multi.jsonl."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402
from rayaccel_b200 import sharding  # noqa: E402

PEAK = 6650.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def main():
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    tris = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    nrays = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    rb.init(local_rank)
    v, i = rb.synthetic_triangles(tris, seed=7, extent=1000.0, edge=2.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scene = rb.create_scene(v, i)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    info = scene.info
    begin, end = sharding.shard_bounds(nrays, rank, world)
    n = end - begin
    # every rank draws the same global ray set chunk by chunk and keeps its slice, so the union over ranks is the
    # 1-GPU ray set whatever the world size
    g = torch.Generator(device="cuda").manual_seed(8)
    lo = torch.tensor(info["bounds_min"], device="cuda")
    hi = torch.tensor(info["bounds_max"], device="cuda")
    rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    chunk = 10_000_000
    for b in range(0, nrays, chunk):
        m = min(chunk, nrays - b)
        o = lo + (hi - lo) * torch.rand((m, 3), device="cuda", generator=g)
        d = torch.randn((m, 3), device="cuda", generator=g)
        d = d / d.norm(dim=1, keepdim=True)
        s, e = max(b, begin), min(b + m, end)
        if s < e:
            rays[s - begin: e - begin, 0:3] = o[s - b: e - b]
            rays[s - begin: e - begin, 4:7] = d[s - b: e - b]
    rays[:, 3] = 0.0
    rays[:, 7] = 1e6
    res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    rb.trace_device(scene, None, [(rays.data_ptr(), res.data_ptr(), n)], counters_ptr=cnt.data_ptr(), detail=True)
    torch.cuda.synchronize()
    if world > 1:
        sharding.reduce_frame_counters(cnt.clone())  # communicator set-up outside the timed part
    stream = torch.cuda.current_stream()
    best = 1e30
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        rb.trace_device(scene, None, [(rays.data_ptr(), res.data_ptr(), n)], stream=stream)
        b.record(stream)
        b.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    sharding.reduce_frame_counters(cnt)
    if rank == 0:
        c = [int(x) for x in cnt.tolist()]
        alg = 48 * c[0] + 64 * c[2] + 48 * c[3] + 4 * c[1] + 64 * (c[0] - c[1])
        line = {"config": f"C5 synthetic soup {tris} triangles, {nrays} uniform random rays, ray-sharded x{world}", "n_gpus": world, "scaling": "strong",
                "rays": c[0], "hit_rate": round(c[1] / c[0], 4), "ms_max_over_ranks": round(best, 3), "mrays": round(c[0] / best / 1e3, 1),
                "alg_bytes_per_ray": round(alg / c[0], 1), "alg_gbs_per_gpu": round(alg / best / 1e6 / world, 1),
                "frac_of_hbm_peak_per_gpu": round(alg / best / 1e6 / world / PEAK, 4), "hbm_peak_gbs": PEAK,
                "scene_build_s_per_rank": round(build_s, 3), "nodes": info["node_count"], "pairs": info["pair_count"],
                "note": "default engine settings (device scene build, re-binning and shared stack tops on: scene >> L2)"}
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/c5_multi.jsonl", "a") as f:
            f.write(json.dumps(line) + "\n")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
