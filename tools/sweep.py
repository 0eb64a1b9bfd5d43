"""Kernel tuning sweep on one GPU: times the traversal launch alone (CUDA events, device-resident
streams) for primary and diffuse-bounce ray sets over a grid of launch shapes. Development tool;
writes JSON lines to gpurun_out/sweep.jsonl."""
import itertools
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402


def time_trace(scene, env, rays, res, n, iters=5, flush=None):
    stream = torch.cuda.current_stream()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        rb.trace_device(scene, env, [(rays.data_ptr(), res.data_ptr(), n)], stream=stream)
        b.record(stream)
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), float(np.median(ts))


def main():
    out_path = os.path.join("gpurun_out", "sweep.jsonl")
    os.makedirs("gpurun_out", exist_ok=True)
    torch.cuda.set_device(0)
    rb.init(0)
    sf = rb.load_scene()
    t0 = time.time()
    scene = rb.create_scene(sf.vertices, sf.indices)
    env = rb.create_environment(sf.environment)
    print("scene build+upload %.3fs" % (time.time() - t0), scene.info)
    w, h, spp = 1920, 1080, int(os.environ.get("SWEEP_SPP", "4"))
    cam = rb.Camera.for_scene(sf, w, h)
    n0 = w * h * spp
    sets = []
    rays = torch.empty(n0 * 8, dtype=torch.float32, device="cuda")
    rb.generate_primary(cam, w, h, spp, 1, rays.data_ptr())
    n = n0
    for bounce in range(4):
        res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
        cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
        rb.trace_device(scene, env, [(rays.data_ptr(), res.data_ptr(), n)], counters_ptr=cnt.data_ptr())
        torch.cuda.synchronize()
        c = cnt.cpu().numpy()
        alg_bytes = 48 * n + 64 * int(c[2]) + 48 * int(c[3]) + 4 * int(c[1]) + 64 * (n - int(c[1]))
        sets.append(dict(name="primary" if bounce == 0 else f"bounce{bounce}", rays=rays, res=res, n=n, alg_bytes=alg_bytes,
                         inner=c[2] / n, pairs=c[3] / n, hit=c[1] / n))
        print(sets[-1]["name"], n, "inner/ray %.2f pairs/ray %.2f hit %.3f bytes/ray %.0f" % (c[2] / n, c[3] / n, c[1] / n, alg_bytes / n))
        nxt = torch.empty(max(int(c[1]), 1) * 8, dtype=torch.float32, device="cuda")
        k = torch.zeros(1, dtype=torch.int32, device="cuda")
        rb.generate_bounce(scene, rays.data_ptr(), res.data_ptr(), n, 2 + bounce, nxt.data_ptr(), k.data_ptr())
        torch.cuda.synchronize()
        rays, n = nxt, int(k.item())
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    grid = []
    grid.append(dict(variant=1))
    for block, smem, thr in itertools.product([128, 256, 512, 1024], [0, -1], [1, 8, 16, 24, 32]):
        grid.append(dict(variant=0, block=block, smem_nodes=smem, fetch_threshold=thr))
    base = dict(variant=0, block=256, ctas_per_sm=0, smem_nodes=-1, fetch_threshold=12)
    with open(out_path, "w") as f:
        for g in grid:
            rb.set_tuning(**{**base, **g})
            row = dict(g)
            tot_t, tot_n, tot_b = 0.0, 0, 0
            for s in sets:
                best, med = time_trace(scene, env, s["rays"], s["res"], s["n"], flush=flush)
                row[s["name"] + "_mrays"] = round(s["n"] / best / 1e3, 1)
                row[s["name"] + "_frac"] = round(s["alg_bytes"] / (best * 1e-3) / 6547.2e9, 4)
                if s["name"] != "primary":
                    tot_t += best; tot_n += s["n"]; tot_b += s["alg_bytes"]
            row["secondary_mrays"] = round(tot_n / tot_t / 1e3, 1)
            row["secondary_frac"] = round(tot_b / (tot_t * 1e-3) / 6547.2e9, 4)
            print(json.dumps(row))
            f.write(json.dumps(row) + "\n")
            f.flush()
    rb.set_tuning(**base)


if __name__ == "__main__":
    main()
