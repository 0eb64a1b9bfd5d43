"""Re-creates one round of tests/fuzz/fuzz_gpu.py from its seed, traces its rays under every kernel variant and writes scene, rays,
images, the checker's results and each variant's results to gpurun_out/fuzz_<seed>.npz for offline analysis."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "fuzz"))
import fuzz_gpu as F  # noqa: E402
import oracle  # noqa: E402
import rayaccel_b200 as rb  # noqa: E402

if torch.cuda.is_available():
    torch.cuda.set_device(0)
rb.init(0)
rb.set_tuning(**F.DEFAULT)
for seed in [int(a) for a in sys.argv[1:]]:
    r = np.random.default_rng(seed)
    kind, verts, indices = F.scene_family(r)
    rb.set_tuning(build_device=int(r.choice([0, 2, 3])))
    scene = rb.create_scene(verts, indices)
    rb.set_tuning(build_device=3)
    env_img = r.random((int(r.integers(1, 9)), int(r.integers(1, 9)), 4)).astype(np.float32) if r.random() < 0.6 else None
    env = rb.create_environment(env_img) if env_img is not None else None
    nodes, pairs, remap = scene.download()
    images = oracle.SceneImages(nodes, pairs, remap, env_img)
    r.random()  # the builder-comparison draw
    rays = F.rays_for(r, verts, indices, int(r.integers(1, 20000)))
    want = F.to_u32(oracle.traverse(images, rays))
    out = dict(verts=verts, indices=indices, nodes=nodes, pairs=pairs, remap=remap, rays=rays, want=want)
    for name, tun in (("v3", dict()), ("v1", dict(variant=1)), ("v2", dict(variant=2)), ("v0", dict(variant=0)), ("v3_smem16", dict(smem_stack=16)), ("v4", dict(variant=4))):
        rb.set_tuning(**{**F.DEFAULT, **tun})
        got = F.trace_device(scene, env, rays)
        rb.set_tuning(**F.DEFAULT)
        bad = np.flatnonzero((got != want).any(1))
        print(seed, kind, name, "differ", bad.size, bad[:6])
        for b in bad[:2]:
            print("   ray", b, rays[b], "got", got[b], got[b].view(np.float32), "want", want[b], want[b].view(np.float32))
        out["got_" + name] = got
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"fuzz_{seed}.npz"), **out)
    scene.destroy()
    print(seed, "variant 4 against north_star's bar:", F.quantised_check(out["got_v4"], want, rays, verts, indices) or "ok")
