"""Development tool: repeated timings of the same launch to see run-to-run spread.  usage: noise.py "k=v,k=v;k=v,..." """
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402

BASE = dict(variant=2, block=256, ctas_per_sm=5, smem_nodes=0, fetch_threshold=16, leaf_bail=4, inner_bail=12, carveout=-1)
torch.cuda.set_device(0)
rb.init(0)
sf = rb.load_scene()
scene = rb.create_scene(sf.vertices, sf.indices)
env = rb.create_environment(sf.environment)
w, h, spp = 1920, 1080, 4
cam = rb.Camera.for_scene(sf, w, h)
n = w * h * spp
rays = torch.empty(n * 8, dtype=torch.float32, device="cuda")
rb.generate_primary(cam, w, h, spp, 1, rays.data_ptr())
sets = []
for bounce in range(4):
    res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    rb.trace_device(scene, env, [(rays.data_ptr(), res.data_ptr(), n)], counters_ptr=cnt.data_ptr(), detail=False)
    torch.cuda.synchronize()
    sets.append((rays, res, n))
    nxt = torch.empty(max(int(cnt[1].item()), 1) * 8, dtype=torch.float32, device="cuda")
    k = torch.zeros(1, dtype=torch.int32, device="cuda")
    rb.generate_bounce(scene, rays.data_ptr(), res.data_ptr(), n, 2 + bounce, nxt.data_ptr(), k.data_ptr())
    torch.cuda.synchronize()
    rays, n = nxt, int(k.item())
descs = [(r.data_ptr(), o.data_ptr(), c) for r, o, c in sets]
stream = torch.cuda.current_stream()
for cfg in (sys.argv[1] if len(sys.argv) > 1 else "variant=0").split(";"):
    t = dict(BASE)
    t.update({k: int(v) for k, v in (p.split("=") for p in cfg.split(","))})
    rb.set_tuning(**t)
    for name, d in (("all", descs), ("sec", descs[1:])):
        ts = []
        for _ in range(12):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            rb.trace_device(scene, env, d, stream=stream)
            b.record(stream)
            b.synchronize()
            ts.append(a.elapsed_time(b))
        print(cfg, name, " ".join(f"{x:.2f}" for x in ts), flush=True)
        for rep in range(3):  # counted launches: does the WORK vary between runs, or only the time?
            cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
            rb.debug_warp_stats(True)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            rb.trace_device(scene, env, d, stream=stream, counters_ptr=cnt.data_ptr(), detail=True)
            b.record(stream)
            b.synchronize()
            st = rb.debug_warp_stats(True)
            print("   counted", f"{a.elapsed_time(b):.2f} ms", "rounds", st[0], "innerIt", st[1], "leafIt", st[2],
                  "lanes/innerIt", round(st[3] / max(st[1], 1), 2), "lanes/leafIt", round(st[4] / max(st[2], 1), 2), "refills", st[5], flush=True)
