"""BASELINE.json config 4: path-tracer wavefront on battlefield.bin, 3840x2160, 16 spp, ray-sharded over the
box's GPUs (strong scaling: the frame is fixed, every rank traces a contiguous slice of it). Development tool.

    python tools/c4_multi.py                                              (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/c4_multi.py

One process per GPU, scene replicated (built on each GPU); a rank owns pixel runs dealt round-robin
(sharding.interleaved_blocks) and traces them, then 3 diffuse bounces of its own hits, one launch per depth. No data-path collective; per frame one NCCL all-reduce of the
frame counters and -- to show its cost -- one all-gather of the primary Result slices. Time = device time
of the traversal launches (CUDA events), max over ranks. Appends one JSON line to gpurun_out/c4.jsonl."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402
from rayaccel_b200 import sharding  # noqa: E402

WIDTH, HEIGHT, SPP, BOUNCES = 3840, 2160, 16, 3


def main():
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else SPP
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    rb.init(local_rank)
    sf = rb.load_scene()
    scene = rb.create_scene(sf.vertices, sf.indices)
    env = rb.create_environment(sf.environment)
    cam = rb.Camera.for_scene(sf, WIDTH, HEIGHT)
    stream = torch.cuda.current_stream()
    pixels = WIDTH * HEIGHT
    # pixel runs dealt round-robin (sky rows are cheap, ground rows expensive: contiguous slices are unbalanced)
    block = int(os.environ.get("C4_BLOCK", "16384"))
    runs = sharding.interleaved_blocks(pixels, rank, world, block) if world > 1 else [(0, pixels)]
    n0 = sum(e - b for b, e in runs)
    frame = torch.zeros(8, dtype=torch.int64, device="cuda")
    trace_ms = 0.0
    gather_ms = 0.0
    rays_traced = 0
    per_depth = [0] * (BOUNCES + 1)

    if world > 1:  # connection set-up of the communicator is not part of a frame
        sharding.reduce_frame_counters(torch.zeros(8, dtype=torch.int64, device="cuda"))
        sharding.gather_results_interleaved(torch.zeros(n0 * 4, dtype=torch.float32, device="cuda"), pixels, block)
        torch.cuda.synchronize()

    def timed_trace(descs, counters):
        descs = rb.pack_streams(descs)  # marshalling thousands of descriptors is Python time, not device time
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        rb.trace_device(scene, env, descs, stream=stream, counters_ptr=counters.data_ptr(), detail=False)
        b.record(stream)
        b.synchronize()
        return a.elapsed_time(b)

    # all samples of the frame at once: sample-major primaries, so the rank's pixels are `spp` contiguous
    # slices = `spp` ray streams traced by ONE launch per depth
    full = torch.empty(spp * pixels * 8, dtype=torch.float32, device="cuda")
    rb.generate_primary(cam, WIDTH, HEIGHT, spp, 1, full.data_ptr(), stream=stream)
    for warm in (True, False):
        streams = [(full[(s * pixels + b) * 8: (s * pixels + e) * 8], e - b) for s in range(spp) for b, e in runs]
        for depth in range(BOUNCES + 1):
            streams = [(r, n) for r, n in streams if n]
            if not streams:
                break
            results = [torch.empty(n * 4, dtype=torch.float32, device="cuda") for _, n in streams]
            cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
            ms = timed_trace([(r.data_ptr(), o.data_ptr(), n) for (r, n), o in zip(streams, results)], cnt)
            if not warm:
                n_all = sum(n for _, n in streams)
                trace_ms += ms
                rays_traced += n_all
                per_depth[depth] += n_all
                frame += cnt
                if depth == 0 and world > 1:
                    sample0 = torch.cat(results[: len(runs)])  # this rank's runs of sample 0, back to back
                    dist.barrier()  # ranks drift apart in the untimed Python parts; do not time the wait for the slowest
                    torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    everything = sharding.gather_results_interleaved(sample0, pixels, block)
                    b.record(stream)
                    b.synchronize()
                    gather_ms = a.elapsed_time(b)
                    assert everything.numel() == pixels * 4
            if depth == BOUNCES:
                break
            # bounce rays of all runs, compacted into one stream per group of runs (one generator call per run)
            nxt_streams = []
            ks = torch.zeros(len(streams), dtype=torch.int32, device="cuda")
            for s, ((r, n), o) in enumerate(zip(streams, results)):
                out = torch.empty(n * 8, dtype=torch.float32, device="cuda")
                rb.generate_bounce(scene, r.data_ptr(), o.data_ptr(), n, 1000 * (s + 1) + depth, out.data_ptr(), ks[s:].data_ptr(), stream=stream)
                nxt_streams.append(out)
            torch.cuda.synchronize()
            streams = [(out, k) for out, k in zip(nxt_streams, ks.tolist())]

    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    sharding.reduce_frame_counters(frame)
    b.record(stream)
    b.synchronize()
    reduce_ms = a.elapsed_time(b)
    t = torch.tensor([trace_ms, gather_ms, reduce_ms], dtype=torch.float64, device="cuda")
    tot = torch.tensor([rays_traced] + per_depth, dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    if rank == 0:
        ms = float(t[0].item())
        total = int(tot[0].item())
        line = {"config": f"C4 path tracer wavefront battlefield {WIDTH}x{HEIGHT} {spp} spp depth {BOUNCES}, ray-sharded x{world}",
                "n_gpus": world, "scaling": "strong", "rays": total, "rays_per_depth": [int(x) for x in tot[1:].tolist()],
                "trace_ms_max_over_ranks": round(ms, 3), "mrays": round(total / ms / 1e3, 1),
                "frame_counter_allreduce_ms": round(float(t[2].item()), 3),
                "primary_result_allgather_ms": round(float(t[1].item()), 3), "allgather_bytes": pixels * 16 if world > 1 else 0,
                "frame_rays_hits": [int(frame[0].item()), int(frame[1].item())],
                "note": "one traversal launch per depth and rank over its 16 sample streams, device time by CUDA events, max over "
                        "ranks; collectives (after a warm-up that sets the communicator up): one all-reduce of the frame counters "
                        "per frame, one all-gather of one sample's primary results"}
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/c4.jsonl", "a") as f:
            f.write(json.dumps(line) + "\n")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
