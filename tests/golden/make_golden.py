#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/. Run in the build container, where
/root/reference exists, after `make -C oracle ref` (the unmodified reference scene builder and
light-probe sampler compiled from their own sources into oracle/_ref/libracc_ref.so):

    python tests/golden/make_golden.py

What is REFERENCE output (produced by executing the reference's own code):
  ref_scene_digests.json  numbering-independent digests of racc::createScene()'s GPU images
                          (Scene.cpp:183-357, Bvh2.cpp) for battlefield.bin and synthetic meshes
  ref_env_samples.npz     racc_internal::sample() (Environment.h:27-82) for 2048 directions
  battlefield_rays.npz    6144 rays (primary, diffuse bounce, uniform random) traced on the REFERENCE-built images
                          by the reference's own traversal kernel SOURCE (Kernels.h:9-242, compiled as C++ over
                          oracle/ref_shim/opencl_c.h into oracle/_ref/libkernel_ref.so by `make -C oracle kernel`
                          and run on the CPU), asserted equal bit for bit to the oracle's results, plus the
                          brute-force fp64 arbiter's (t, id)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import raygen  # noqa: E402
from rayaccel_b200 import scene_io  # noqa: E402


def synthetic_meshes():
    """Name -> (verts4, indices): the meshes whose reference digests are pinned."""
    out = {}
    for n, seed in ((3, 1), (4, 2), (33, 3), (1000, 4), (20000, 5)):
        out[f"soup_{n}"] = scene_io.synthetic_triangles(n, seed=seed, extent=100.0, edge=3.0)
    g = 40  # regular grid: many equal centroids -> exercises the builder's tie-breaking
    xs, zs = np.meshgrid(np.arange(g + 1, dtype=np.float32), np.arange(g + 1, dtype=np.float32))
    v = np.ones(((g + 1) * (g + 1), 4), np.float32)
    v[:, 0], v[:, 1], v[:, 2] = xs.ravel(), 0.0, zs.ravel()
    a = (np.arange(g)[:, None] * (g + 1) + np.arange(g)[None, :]).ravel().astype(np.uint32)
    idx = np.stack([a, a + g + 1, a + 1, a + 1, a + g + 1, a + g + 2], axis=1).ravel().astype(np.uint32)
    out["grid_40"] = (v, idx)
    return out


def main():
    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    sf = scene_io.load_scene()
    digests = {}
    ref_img = oracle.ref_build_scene(sf.vertices, sf.indices)
    digests["battlefield"] = ref_img.digest()
    for name, (v, i) in synthetic_meshes().items():
        digests[name] = oracle.ref_build_scene(v, i).digest()
    with open(os.path.join(HERE, "ref_scene_digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)

    rng = np.random.default_rng(11)
    dirs = rng.normal(size=(2048, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    dirs[:6] = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float32)
    np.savez_compressed(os.path.join(HERE, "ref_env_samples.npz"), dirs=dirs, rgb=oracle.ref_env_sample(sf.environment, dirs))

    ref_img.env = sf.environment
    cam = raygen.look_at(sf.cam_origin, sf.cam_target, sf.cam_up, sf.cam_fov, 1920, 1080)
    rows = np.linspace(0, 1079, 16).astype(np.int64)
    primary_all = raygen.primary_rays(cam, 1920, 1080, 1, 0, rows=rows)
    primary = primary_all[:: primary_all.shape[0] // 2048][:2048]
    res = oracle.traverse(ref_img, primary_all)
    bounce = raygen.bounce_rays(sf.vertices, sf.indices, primary_all, res, seed=2)
    bounce = bounce[:: max(1, bounce.shape[0] // 2048)][:2048]
    lo, hi = sf.vertices[:, :3].min(0), sf.vertices[:, :3].max(0)
    rnd = np.zeros(2048, oracle.RAY_DTYPE)
    rnd["origin"] = rng.uniform(lo, hi, size=(2048, 3)).astype(np.float32)
    d = rng.normal(size=(2048, 3))
    rnd["dir"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rnd["minT"], rnd["maxT"] = 0.0, 1e6
    rays = np.concatenate([primary, bounce, rnd])
    results, counters = oracle.traverse(ref_img, rays, counters=True)
    # `results` are the outputs of the REFERENCE ITSELF: its unmodified builder's images walked by its own traversal
    # kernel, compiled from the source text of Kernels.h (oracle/_ref/libkernel_ref.so). The oracle must agree bit for bit.
    assert oracle.have_ref_kernel(), "build oracle/_ref first (make -C oracle kernel)"
    ref_results = oracle.ref_kernel_traverse(ref_img, rays)
    assert np.array_equal(ref_results.view(np.uint32), results.view(np.uint32)), "oracle differs from the reference kernel"
    t64, id64 = oracle.brute_f64(sf.vertices, sf.indices, rays)
    np.savez_compressed(os.path.join(HERE, "battlefield_rays.npz"), rays=rays.view(np.float32).reshape(-1, 8),
                        results=ref_results.view(np.uint32).reshape(-1, 4), inner=counters["inner"], pairs=counters["pairs"],
                        t64=t64, id64=id64, kinds=np.array([primary.shape[0], bounce.shape[0], rnd.shape[0]]),
                        results_by=np.array("reference builder + reference traversal kernel source (oracle/_ref), run on the CPU"))
    for fn in sorted(os.listdir(HERE)):
        print(f"{fn}: {os.path.getsize(os.path.join(HERE, fn))} bytes")


if __name__ == "__main__":
    main()
