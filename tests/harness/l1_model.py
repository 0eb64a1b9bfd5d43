#!/usr/bin/env python
"""L1 gather model of the default traversal kernel, evaluated WITHOUT a GPU.

tracePackedKernel is bound by the L1 data pipe: what a ray costs is the number of data-pipe
wavefronts its node and pair gathers need, and that depends on how many lanes of a warp fetch the same sectors in the
same instruction. This tool runs the kernel's own source on the CPU (tests/harness/cuda_on_cpu),
traces every 256-bit gather per warp-level instruction instance, and prices it with the rule measured on the B200
(profiles/r01_l1_wavefront_microbench.md): max(0.266 x lanes, 1.065 x distinct 32-byte sectors) wavefronts, one
wavefront per clock per SM, 148 SMs x 1.965 GHz = 291 G wavefronts/s. It answers "would this visiting order / refill
policy need fewer wavefronts per ray?" before any GPU minute is spent; it says nothing about issue slots (what bounds
coherent primaries), latency or DRAM.

One thing differs from the device: there 148 x 5 x 8 = 5920 warps pull from the cursor at once, so successive refills of
ONE warp are ~95 K rays apart; here one CTA runs at a time. `--spread` (default) visits the rays through a permutation
that hands consecutive 16-ray chunks out 5920 chunks apart, which reproduces that; `--no-spread` is arrival order.

    python tests/harness/l1_model.py --width 480 --height 270 --spp 1

Test infrastructure (it lives under tests/ because it builds on tests/harness/cuda_on_cpu and uses the checker).
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402  (test infrastructure: the checker supplies the hits the bounce generator needs)
import rayaccel_b200 as rb  # noqa: E402
from oracle import raygen  # noqa: E402
from test_kernels_on_cpu import build_tracer, trace_on_cpu  # noqa: E402

PEAK_WAVEFRONTS = 148 * 1.965e9


def spread_permutation(n, chunk=16, gpu_warps=5920, emulated_warps=8):
    """Position p of the visiting order -> ray index. On the device, refill slot s of the cursor (a chunk of ~16 rays) goes
    to warp s % 5920 as its (s // 5920)-th refill; here the CTA's 8 warps take slots in turn. Emulated warp e therefore
    plays device warps e, e + 8, e + 16, ... one after the other, each for its R = slots / 5920 refills: its k-th slot is
    device slot  w + (k % R) * 5920  with  w = e + 8 * (k // R)."""
    slots = n // chunk
    refills = max(slots // gpu_warps, 1)
    covered = min(slots, refills * gpu_warps)
    j = np.arange(covered)
    e, k = j % emulated_warps, j // emulated_warps
    w = e + emulated_warps * (k // refills)
    s = w + (k % refills) * gpu_warps
    ok = (w < gpu_warps) & (s < covered)
    used = np.zeros(slots, bool)
    used[s[ok]] = True
    order = np.concatenate([s[ok], np.flatnonzero(~used)])  # whatever the mapping leaves out comes last, in order
    idx = (order[:, None] * chunk + np.arange(chunk)[None, :]).reshape(-1)
    idx = np.concatenate([idx, np.arange(slots * chunk, n)])
    assert len(idx) == n and len(np.unique(idx)) == n
    return idx.astype(np.uint32)


def morton_permutation(rays, lo, hi, bits=5):
    q = np.clip(((rays["origin"] - lo) / np.maximum(hi - lo, 1e-9) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    key = np.zeros(len(rays), np.int64)
    for b in range(bits):
        for a in range(3):
            key |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return np.argsort(key, kind="stable").astype(np.uint32)


def model(lib, images, rays, perm, tuning):
    out = (ctypes.c_double * 6)()
    lib.cpu_l1_model_begin()
    t0 = time.time()
    trace_on_cpu(lib, images, [rays], perm=perm, tuning=tuning)
    lib.cpu_l1_model_end(out)
    n = len(rays)
    return {"rays": n, "gathers_per_ray": round(out[1] / n, 2), "lanes_per_instruction": round(out[1] / max(out[2], 1), 2),
            "wavefronts_per_ray_by_sharing_group": {"32": round(out[0] / n, 2), "8": round(out[3] / n, 2), "4": round(out[4] / n, 2),
                                                    "2": round(out[5] / n, 2), "1": round(1.065 * out[1] / n, 2)},
            "gray_per_s_at_80pct_of_the_pipe": {k: round(0.8 * PEAK_WAVEFRONTS / (w / n) / 1e9, 2)
                                                for k, w in (("32", out[0]), ("8", out[3]), ("4", out[4]), ("2", out[5]), ("1", 1.065 * out[1]))},
            "seconds": round(time.time() - t0, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=480)
    ap.add_argument("--height", type=int, default=270)
    ap.add_argument("--spp", type=int, default=1)
    ap.add_argument("--bounces", type=int, default=2)
    ap.add_argument("--no-spread", action="store_true")
    ap.add_argument("--tuning", default="256,5,16,4,8,0", help="block, CTAs/SM, refill threshold, leaf bail, inner bail, smem stack")
    args = ap.parse_args()
    tuning = tuple(int(x) for x in args.tuning.split(","))
    lib = ctypes.CDLL(build_tracer())
    lib.cpu_l1_model_end.argtypes = [ctypes.POINTER(ctypes.c_double)]
    sf = rb.load_scene()
    hi_ = rb.HostImages(sf.vertices, sf.indices)
    images = oracle.SceneImages(hi_.nodes, hi_.pairs, hi_.remap, sf.environment)
    cam = raygen.look_at(sf.cam_origin, sf.cam_target, sf.cam_up, sf.cam_fov, args.width, args.height)
    lo, hi = sf.vertices[:, :3].min(0), sf.vertices[:, :3].max(0)
    rays = raygen.primary_rays(cam, args.width, args.height, spp=args.spp, seed=1)
    for wave in range(args.bounces + 1):
        name = "primary" if wave == 0 else f"bounce{wave}"
        orders = {"arrival": None if args.no_spread else spread_permutation(len(rays))}
        m = morton_permutation(rays, lo, hi)
        orders["rebinned (origin Morton, 5 bits/axis)"] = m if args.no_spread else m[spread_permutation(len(rays))]
        for label, perm in orders.items():
            print(json.dumps({"stream": name, "order": label, "spread": not args.no_spread, "tuning": list(tuning), **model(lib, images, rays, perm, tuning)}), flush=True)
        res = oracle.traverse(images, rays)
        rays = raygen.bounce_rays(sf.vertices, sf.indices, rays, res, seed=2 + wave)
        if len(rays) == 0:
            break


if __name__ == "__main__":
    main()
