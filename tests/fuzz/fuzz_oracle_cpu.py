"""CPU-only campaign: the checker (oracle/racc_oracle.c) against the reference's OWN kernel source (Kernels.h compiled over
oracle/ref_shim/opencl_c.h into oracle/_ref/libkernel_ref.so) on the random scene families and adversarial rays of
tests/fuzz/fuzz_gpu.py: triangle id, t, u, v and miss radiance bit for bit. This is what pins the checker. Needs oracle/_ref.

    python tests/fuzz/fuzz_oracle_cpu.py [--seconds 120] [--seed 1]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "fuzz"))
import oracle  # noqa: E402
import rayaccel_b200 as rb  # noqa: E402
from fuzz_gpu import rays_for, scene_family  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    assert oracle.have_ref_kernel(), "oracle/_ref/libkernel_ref.so is not built (needs /root/reference)"
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    rounds = rays_total = 0
    kinds, failures = {}, []
    while time.time() - t0 < args.seconds and len(failures) < 5:
        seed = int(rng.integers(0, 2 ** 31))
        r = np.random.default_rng(seed)
        kind, verts, indices = scene_family(r)
        h = rb.HostImages(verts, indices)
        env = r.random((int(r.integers(1, 9)), int(r.integers(1, 9)), 4)).astype(np.float32) if r.random() < 0.6 else None
        images = oracle.SceneImages(h.nodes, h.pairs, h.remap, env)
        rays = rays_for(r, verts, indices, int(r.integers(1, 20000)))
        a = oracle.traverse(images, rays).view(np.uint32).reshape(-1, 4)
        b = oracle.ref_kernel_traverse(images, rays, threads=0).view(np.uint32).reshape(-1, 4)
        if not np.array_equal(a, b):
            bad = np.flatnonzero((a != b).any(1))
            failures.append(f"round {rounds} seed {seed} {kind} {len(indices) // 3} triangles: {bad.size}/{len(rays)} rays differ, first {bad[0]}: checker {a[bad[0]]} reference kernel {b[bad[0]]}")
        rays_total += len(rays)
        kinds[kind] = kinds.get(kind, 0) + 1
        rounds += 1
    print(f"checker vs reference kernel source: {rounds} scenes, {rays_total} rays in {time.time() - t0:.0f} s, families { {str(k): v for k, v in kinds.items()} }")
    for f in failures:
        print("FAIL", f)
    print("ok" if not failures else f"{len(failures)} failures")
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
