"""BASELINE configs[4] for ncu: the 10 M-triangle soup and the 100 M uniform random rays of bench.py's `c5` record (same
seeds), one warm-up pass, then one pass with the exact node format (variant 3) and one with the quantised one (variant 4).
usage: prof_c5.py [triangles [rays]]   (run under: ncu --set full -k regex:tracePackedKernel -s 1 -c 2 ...)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402

tris = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
torch.cuda.set_device(0)
rb.init(0)
v, i = rb.synthetic_triangles(tris, seed=7, extent=1000.0, edge=2.0)
scene = rb.create_scene(v, i)
g = torch.Generator(device="cuda")
g.manual_seed(8)
rays = torch.empty(n, 8, dtype=torch.float32, device="cuda")
rays[:, 0:3] = torch.rand(n, 3, generator=g, device="cuda") * 1000.0
d = torch.randn(n, 3, generator=g, device="cuda")
rays[:, 4:7] = d / d.norm(dim=1, keepdim=True)
rays[:, 3] = 0.0
rays[:, 7] = 1e6
del d
res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
desc = [(rays.data_ptr(), res.data_ptr(), n)]
rb.trace_device(scene, None, desc)  # warm-up (skipped by ncu -s 1)
for tuning in (dict(variant=3), dict(variant=4)):
    variant = tuning["variant"]
    rb.set_tuning(**tuning)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rb.trace_device(scene, None, desc)
    b.record()
    b.synchronize()
    print(f"{tuning}: {a.elapsed_time(b):.3f} ms, {n / a.elapsed_time(b) / 1e3:.1f} Mrays/s (sort + traversal)")
rb.set_tuning(variant=3)
