"""Per-stream timing of the bench batch (primary, bounce1..3 launched separately) under a list of tunings.
Development tool: does launch size interact with the launch shape?  usage: sweep_streams.py 'k=v,k=v;k=v'"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402
from sweep import BASE, time_launch  # noqa: E402


def main():
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
    for bounce in range(6):
        res = torch.empty(n * 4, dtype=torch.float32, device="cuda")
        cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
        rb.trace_device(scene, env, [(rays.data_ptr(), res.data_ptr(), n)], counters_ptr=cnt.data_ptr(), detail=True)
        torch.cuda.synchronize()
        c = [int(x) for x in cnt.tolist()]
        sets.append((rays, res, n, round(c[2] / n, 1), round(c[3] / n, 1)))
        nxt = torch.empty(max(c[1], 1) * 8, dtype=torch.float32, device="cuda")
        k = torch.zeros(1, dtype=torch.int32, device="cuda")
        rb.generate_bounce(scene, rays.data_ptr(), res.data_ptr(), n, 2 + bounce, nxt.data_ptr(), k.data_ptr())
        torch.cuda.synchronize()
        rays, n = nxt, int(k.item())
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
    print("streams:", [(s[2], s[3], s[4]) for s in sets])
    for spec in sys.argv[1].split(";"):
        g = {k: int(v) for k, v in (p.split("=") for p in spec.split(",") if p)}
        rb.set_tuning(**{**BASE, **g})
        row = dict(g)
        for i, (r, o, c, _, _) in enumerate(sets):
            t = time_launch(scene, env, [(r.data_ptr(), o.data_ptr(), c)], iters=5, flush=flush)
            row[f"s{i}_mrays"] = round(c / t / 1e3, 1)
        print(json.dumps(row), flush=True)
    rb.set_tuning(**BASE)


if __name__ == "__main__":
    main()
