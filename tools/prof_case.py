"""Launches the bench batch's primary stream and its three bounce streams as two separate traversal
launches under a given tuning, for ncu.  usage: prof_case.py key=value,key=value [repeat]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402

tuning = dict(variant=3, block=256, ctas_per_sm=5, smem_nodes=0, fetch_threshold=16, leaf_bail=4, inner_bail=8, carveout=-1,
              sort=0, sort_origin_bits=5, sort_dir_bits=0, sort_dir_major=0, smem_stack=0)
if len(sys.argv) > 1 and sys.argv[1]:
    tuning.update({k: int(v) for k, v in (p.split("=") for p in sys.argv[1].split(","))})
repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 2
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
rb.set_tuning(variant=1)  # build the streams with the simple kernel so that ncu -k regex:tracePersistent skips them
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
rb.set_tuning(**tuning)
for _ in range(repeat):
    rb.trace_device(scene, env, descs[:1])
    rb.trace_device(scene, env, descs[1:])
torch.cuda.synchronize()
print("done", tuning)
