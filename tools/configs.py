"""Measures the BASELINE.json configurations that are not the bench line (development tool, one GPU):

    python tools/configs.py c3 [spp=64] [bounces=8]     path-tracer wavefront, 1920x1080, incoherent stress
    python tools/configs.py c5 [tris=10000000] [rays=100000000]   synthetic soup + uniform random rays

Device-resident streams, CUDA events around the traversal launches only, one JSON line per config
appended to gpurun_out/configs.jsonl. Algorithmic bytes per SURVEY.md section 8d from the kernel's own
visit counters (a separate, untimed counted pass)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def alg_bytes(n, c):
    return 48 * n + 64 * c[2] + 48 * c[3] + 4 * c[1] + 64 * (n - c[1])


def timed_trace(scene, env, descs, iters=3):
    stream = torch.cuda.current_stream()
    best = 1e30
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        rb.trace_device(scene, env, descs, stream=stream)
        b.record(stream)
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def counted(scene, env, rays, res, n):
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    rb.trace_device(scene, env, [(rays.data_ptr(), res.data_ptr(), n)], counters_ptr=cnt.data_ptr(), detail=True)
    torch.cuda.synchronize()
    return [int(x) for x in cnt.cpu().tolist()]


def emit(line):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/configs.jsonl", "a") as f:
        f.write(json.dumps(line) + "\n")
    print(json.dumps(line), flush=True)


def c3(spp_total=64, bounces=8, width=1920, height=1080, spp_batch=4):
    sf = rb.load_scene()
    scene = rb.create_scene(sf.vertices, sf.indices)
    env = rb.create_environment(sf.environment)
    cam = rb.Camera.for_scene(sf, width, height)
    per_depth = [dict(rays=0, ms=0.0, bytes=0, hits=0) for _ in range(bounces + 1)]
    for batch in range(spp_total // spp_batch):
        n = width * height * spp_batch
        rays = torch.empty(n * 8, dtype=torch.float32, device="cuda")
        rb.generate_primary(cam, width, height, spp_batch, 1 + batch, rays.data_ptr())
        for depth in range(bounces + 1):
            if n == 0:
                break
            res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
            c = counted(scene, env, rays, res, n)
            ms = timed_trace(scene, env, [(rays.data_ptr(), res.data_ptr(), n)])
            d = per_depth[depth]
            d["rays"] += n; d["ms"] += ms; d["bytes"] += alg_bytes(n, c); d["hits"] += c[1]
            if depth == bounces:
                break
            nxt = torch.empty(max(c[1], 1) * 8, dtype=torch.float32, device="cuda")
            k = torch.zeros(1, dtype=torch.int32, device="cuda")
            rb.generate_bounce(scene, rays.data_ptr(), res.data_ptr(), n, 1000 * (batch + 1) + depth, nxt.data_ptr(), k.data_ptr())
            torch.cuda.synchronize()
            rays, n = nxt, int(k.item())

    def agg(ds):
        r, ms, b = sum(d["rays"] for d in ds), sum(d["ms"] for d in ds), sum(d["bytes"] for d in ds)
        return dict(rays=r, ms=round(ms, 3), mrays=round(r / ms / 1e3, 1) if ms else 0, alg_gbs=round(b / ms / 1e6, 1) if ms else 0,
                    frac_of_measured_hbm=round(b / ms / 1e6 / PEAK, 4) if ms else 0)
    emit({"config": f"C3 path tracer wavefront battlefield {width}x{height} {spp_total} spp {bounces} bounces (one launch per depth per {spp_batch}-spp batch)",
          "primary": agg(per_depth[:1]), "secondary_bounce_ge1": agg(per_depth[1:]), "secondary_bounce_ge2": agg(per_depth[2:]), "all": agg(per_depth),
          "per_depth": [dict(depth=i, rays=d["rays"], hit_rate=round(d["hits"] / max(d["rays"], 1), 4), mrays=round(d["rays"] / d["ms"] / 1e3, 1) if d["ms"] else 0)
                        for i, d in enumerate(per_depth)], "hbm_peak_gbs": PEAK})


def c5(tris=10_000_000, nrays=100_000_000):
    t0 = time.perf_counter()
    v, i = rb.synthetic_triangles(tris, seed=7, extent=1000.0, edge=2.0)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    scene = rb.create_scene(v, i)
    t_build = time.perf_counter() - t0
    info = scene.info
    g = torch.Generator(device="cuda").manual_seed(8)
    lo = torch.tensor(info["bounds_min"], device="cuda")
    hi = torch.tensor(info["bounds_max"], device="cuda")
    rays = torch.empty((nrays, 8), dtype=torch.float32, device="cuda")
    chunk = 10_000_000
    for b in range(0, nrays, chunk):
        m = min(chunk, nrays - b)
        rays[b:b + m, 0:3] = lo + (hi - lo) * torch.rand((m, 3), device="cuda", generator=g)
        d = torch.randn((m, 3), device="cuda", generator=g)
        rays[b:b + m, 4:7] = d / d.norm(dim=1, keepdim=True)
    rays[:, 3] = 0.0
    rays[:, 7] = 1e6
    res = torch.empty(nrays * 4, dtype=torch.float32, device="cuda")
    c = counted(scene, None, rays, res, nrays)
    b = alg_bytes(nrays, c)
    scene_mb = (info["node_count"] * 64 + info["pair_count"] * 48 + info["remap_count"] * 4) / 1e6
    ref = None
    # the checker's answer for a strided sample of >= 1 M rays at FULL scene size (not `same_bits_as_first` alone)
    import oracle  # noqa: E402  (test infrastructure; this is a development tool, not the product)
    nodes, pairs, remap = scene.download()
    img = oracle.SceneImages(nodes, pairs, remap)
    stride = max(1, nrays // 1_000_000)
    idx = torch.arange(0, nrays, stride, device="cuda")
    sample = rays[idx].cpu().numpy().reshape(-1).view(oracle.RAY_DTYPE)
    want = torch.from_numpy(oracle.traverse(img, sample).view(np.int32).reshape(-1, 4).copy()).cuda()
    del img, nodes, pairs, remap
    # RACC_CFG_TUNINGS="k=v,k=v;k=v": time the same rays under several tunings (one line each)
    for spec in os.environ.get("RACC_CFG_TUNINGS", "").split(";"):
        tuning = {k: int(v) for k, v in (kv.split("=") for kv in spec.split(",") if kv)}
        rb.set_tuning(**tuning)
        ms = timed_trace(scene, None, [(rays.data_ptr(), res.data_ptr(), nrays)])
        same = None
        if ref is None:
            ref = res.clone()
        else:
            same = bool(torch.equal(ref.view(torch.int32), res.view(torch.int32)))
        got = res.view(-1, 4).view(torch.int32)[idx]
        emit({"config": f"C5 synthetic soup {tris} triangles, {nrays} uniform random rays", "tuning": tuning, "same_bits_as_first": same,
              "parity_sample_bit_exact": bool(torch.equal(got, want)), "parity_sample_ids_differing": int((got[:, 0] != want[:, 0]).sum()),
              "parity_sample_rays": int(idx.numel()),
              "rays": nrays, "ms": round(ms, 3), "mrays": round(nrays / ms / 1e3, 1),
              "hit_rate": round(c[1] / nrays, 4), "inner_per_ray": round(c[2] / nrays, 2), "pairs_per_ray": round(c[3] / nrays, 2),
              "alg_bytes_per_ray": round(b / nrays, 1), "alg_gbs": round(b / ms / 1e6, 1), "frac_of_measured_hbm": round(b / ms / 1e6 / PEAK, 4),
              "scene_mb": round(scene_mb, 1), "nodes": info["node_count"], "pairs": info["pair_count"], "depth": info["depth"],
              "host_build_s": round(t_build, 2), "mesh_gen_s": round(t_gen, 2), "hbm_peak_gbs": PEAK})


if __name__ == "__main__":
    torch.cuda.set_device(0)
    rb.init(0)
    which = sys.argv[1] if len(sys.argv) > 1 else "c5"
    args = [int(a) for a in sys.argv[2:]]
    if which == "c3":
        c3(*args)
    else:
        c5(*args)
