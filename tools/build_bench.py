"""Scene-build timing, host threads vs GPU (bvh_build.cu), same tree. Development tool (one GPU).
usage: build_bench.py [triangle counts...]   -> JSON lines in gpurun_out/build_bench.jsonl"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402


def timed(build_device, v, i, repeat=2):
    rb.set_tuning(build_device=build_device)
    best, scene = 1e30, None
    for _ in range(repeat):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        scene = rb.create_scene(v, i)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    rb.set_tuning(build_device=3)
    return best, scene


def main():
    torch.cuda.set_device(0)
    rb.init(0)
    os.makedirs("gpurun_out", exist_ok=True)
    cases = [("battlefield", None)] + [(f"soup {n}", n) for n in (int(a) for a in sys.argv[1:])]
    with open("gpurun_out/build_bench.jsonl", "a") as f:
        for name, n in cases:
            if n is None:
                sf = rb.load_scene()
                v, i = sf.vertices, sf.indices
            else:
                v, i = rb.synthetic_triangles(n, seed=7, extent=1000.0, edge=2.0)
            th, sh = timed(0, v, i)
            t1, s1 = timed(1, v, i)
            t2, s2 = timed(2, v, i)
            a = sh.download()
            same = all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for sd in (s1, s2) for x, y in zip(a, sd.download())) \
                and sh.info == s1.info == s2.info
            line = {"scene": name, "triangles": int(len(i) // 3), "host_build_s": round(th, 4), "device_sah_host_packing_s": round(t1, 4),
                    "device_build_s": round(t2, 4), "speedup": round(th / t2, 2), "images_identical": bool(same),
                    "nodes": sh.info["node_count"], "depth": sh.info["depth"],
                    "note": "create_scene wall time (input on the host, images ready on the device, incl. the packed copies): host = "
                            "SAH + order + pair merge + packing on host threads, then upload; device = everything on the GPU"}
            print(json.dumps(line), flush=True)
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
