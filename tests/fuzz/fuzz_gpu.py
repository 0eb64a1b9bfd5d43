"""Randomised parity campaign on a GPU: many small random scenes x adversarial rays, engine against the CPU checker, bit for bit.

Each round draws a scene family (soup, shared-edge mesh, coincident copies, slivers and zero-area triangles, axis-aligned
quads, huge / tiny / far-from-origin coordinates), builds it through racc_cuda_scene_create (host or device builder, drawn
per round; a few rounds compare the two builders' images), draws rays that aim at vertices, edge midpoints and triangle
interiors from inside and outside the bounds, start ON triangles, run along axes, carry empty or tiny [minT, maxT], and traces
them (a) as one DEVICE stream under the default tuning, (b) under a second tuning drawn from the launch shapes the tests
use, (c) as ragged HOST streams. All three must equal oracle.traverse on the images the kernel walks, as 32-bit words.
A few rounds render a small frame with racc_cuda_path_trace (both forms) against oracle.path_trace.

    python tests/fuzz/fuzz_gpu.py [--seconds 240] [--seed 1]      # development tool; log copied to profiles/ by hand
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (the checker; this is a test tool)
import rayaccel_b200 as rb  # noqa: E402

DEFAULT = dict(variant=3, block=256, ctas_per_sm=5, smem_nodes=0, fetch_threshold=16, inner_bail=8, leaf_bail=4,
               sort=0, sort_origin_bits=5, sort_dir_bits=0, sort_dir_major=0, smem_stack=0)
SHAPES = [
    dict(variant=0, block=256, smem_nodes=-1, ctas_per_sm=0), dict(variant=1), dict(variant=2),
    dict(variant=2, inner_bail=32, leaf_bail=32, fetch_threshold=1),
    dict(variant=3, ctas_per_sm=4, fetch_threshold=1, inner_bail=32, leaf_bail=32),
    dict(variant=3, block=128, ctas_per_sm=10, inner_bail=0, leaf_bail=0),
    dict(variant=3, block=512, ctas_per_sm=2, inner_bail=20, leaf_bail=1, fetch_threshold=32),
    dict(variant=3, smem_stack=16), dict(variant=3, smem_stack=8, fetch_threshold=4),
    dict(variant=3, sort=1), dict(variant=3, sort=1, sort_origin_bits=6, sort_dir_bits=4, sort_dir_major=1, fetch_threshold=8),
]


BIG_SHARE = 0.02  # --big raises it


def scene_family(rng):
    kind = rng.choice(["soup", "mesh", "coincident", "slivers", "quads", "huge", "tiny", "offset", "few"])
    scale, offset = 1.0, np.zeros(3)
    big = rng.random() < BIG_SHARE  # now and then a scene of the size where the device builder's other code paths run
    if kind == "soup":
        n = int(rng.integers(3, 4000)) if not big else int(rng.integers(20000, 250000))
        c = rng.uniform(-50, 50, (n, 1, 3))
        v = (c + rng.normal(0, rng.uniform(0.05, 5.0), (n, 3, 3))).reshape(-1, 3)
        i = np.arange(3 * n, dtype=np.uint32)
    elif kind == "mesh":  # height field: shared vertices and edges, neighbouring triangles become pairs
        w, h = (int(rng.integers(2, 60)), int(rng.integers(2, 60))) if not big else (int(rng.integers(100, 350)), int(rng.integers(100, 350)))
        xs, ys = np.meshgrid(np.arange(w + 1, dtype=np.float64), np.arange(h + 1, dtype=np.float64))
        z = rng.normal(0, rng.uniform(0.0, 2.0), xs.shape)
        v = np.stack([xs, z, ys], -1).reshape(-1, 3)
        q = (np.arange(h)[:, None] * (w + 1) + np.arange(w)[None, :]).reshape(-1)
        i = np.stack([q, q + 1, q + w + 2, q, q + w + 2, q + w + 1], -1).reshape(-1).astype(np.uint32)
    elif kind == "coincident":  # the same triangles several times: exact ties in t
        n = int(rng.integers(1, 300))
        base = rng.uniform(-10, 10, (n, 3, 3))
        v = np.concatenate([base] * int(rng.integers(2, 5))).reshape(-1, 3)
        i = np.arange(v.shape[0], dtype=np.uint32)
    elif kind == "slivers":  # long thin, collinear and zero-area triangles among ordinary ones
        n = int(rng.integers(4, 1500))
        a = rng.uniform(-20, 20, (n, 3))
        d = rng.normal(size=(n, 3))
        b = a + d * rng.uniform(0.0, 40.0, (n, 1))
        c = a + d * rng.uniform(0.0, 40.0, (n, 1)) + rng.normal(0, 1e-4, (n, 3)) * (rng.random((n, 1)) < 0.5)
        deg = rng.random(n) < 0.1
        c[deg] = a[deg]
        v = np.stack([a, b, c], 1).reshape(-1, 3)
        i = np.arange(3 * n, dtype=np.uint32)
    elif kind == "quads":  # axis-aligned quads: boxes of zero thickness
        n = int(rng.integers(1, 800))
        o = np.round(rng.uniform(-30, 30, (n, 3)))
        ax = rng.integers(0, 3, n)
        e1 = np.zeros((n, 3)); e2 = np.zeros((n, 3))
        e1[np.arange(n), (ax + 1) % 3] = rng.integers(1, 6, n)
        e2[np.arange(n), (ax + 2) % 3] = rng.integers(1, 6, n)
        v = np.stack([o, o + e1, o + e1 + e2, o, o + e1 + e2, o + e2], 1).reshape(-1, 3)
        i = np.arange(6 * n, dtype=np.uint32)
    elif kind in ("huge", "tiny", "offset"):
        n = int(rng.integers(3, 2000))
        c = rng.uniform(-50, 50, (n, 1, 3))
        v = (c + rng.normal(0, 1.5, (n, 3, 3))).reshape(-1, 3)
        i = np.arange(3 * n, dtype=np.uint32)
        if kind == "huge": scale = 10.0 ** rng.uniform(3, 9)
        if kind == "tiny": scale = 10.0 ** rng.uniform(-9, -3)
        if kind == "offset": offset = rng.uniform(-1, 1, 3) * 10.0 ** rng.uniform(3, 6)
    else:  # "few": 1..4 triangles (root-leaf scenes)
        n = int(rng.integers(1, 5))
        v = rng.uniform(-5, 5, (3 * n, 3))
        i = np.arange(3 * n, dtype=np.uint32)
    v = (v * scale + offset).astype(np.float32)
    verts = np.zeros((v.shape[0], 4), np.float32)
    verts[:, :3] = v
    return kind, verts, i


def rays_for(rng, verts, indices, n):
    v = verts[:, :3].astype(np.float64)
    tri = v[indices.reshape(-1, 3).astype(np.int64)]
    lo, hi = v.min(0), v.max(0)
    ext = np.maximum(hi - lo, 1e-30)
    pick = tri[rng.integers(0, tri.shape[0], n)]
    bary = rng.dirichlet([1, 1, 1], n)
    mode = rng.integers(0, 6, n)
    target = (pick * bary[:, :, None]).sum(1)                                   # interior points
    target[mode == 1] = pick[mode == 1, rng.integers(0, 3)]                     # a vertex, exactly
    target[mode == 2] = 0.5 * (pick[mode == 2, 0] + pick[mode == 2, 1])         # an edge midpoint
    origin = lo + rng.uniform(-0.5, 1.5, (n, 3)) * ext                          # inside and around the bounds
    on = mode == 3                                                              # starts ON a triangle
    origin[on] = target[on]
    target[on] = lo + rng.uniform(0, 1, (int(on.sum()), 3)) * ext
    d = target - origin
    rnd = mode >= 4
    d[rnd] = rng.normal(size=(int(rnd.sum()), 3))
    ln = np.linalg.norm(d, axis=1, keepdims=True)
    d = np.where(ln > 0, d / np.maximum(ln, 1e-300), [[1.0, 0.0, 0.0]])
    axis = rng.random(n) < 0.08                                                 # axis-parallel: zero components
    k = rng.integers(0, 3, n)
    d[axis] = 0.0
    d[axis, k[axis]] = rng.choice([-1.0, 1.0], int(axis.sum()))
    r = np.zeros(n, dtype=oracle.RAY_DTYPE)
    r["origin"], r["dir"] = origin.astype(np.float32), d.astype(np.float32)
    r["minT"] = np.where(rng.random(n) < 0.2, rng.choice([1e-3, 1e-6, 1.0]) * float(ext.max()), 0.0).astype(np.float32)
    far = np.float32(1e6) * np.float32(max(1.0, float(ext.max())))
    r["maxT"] = far
    short = rng.random(n) < 0.1
    r["maxT"][short] = (np.linalg.norm(ln[short], axis=1) * rng.uniform(0.3, 1.3, int(short.sum()))).astype(np.float32)
    empty = rng.random(n) < 0.02
    r["maxT"][empty] = r["minT"][empty]
    return r


def quantised_check(got, want, rays, verts, indices):
    """Variant 4 walks conservative boxes: same triangle -> same bits; a different triangle must be a closest hit by the fp64
    brute force (a tie, or a hit the exact fp32 boxes let slip); a miss where the checker hits is never allowed."""
    same = got[:, 0] == want[:, 0]
    if not np.array_equal(got[same], want[same]):
        return "same triangle (or both miss) but different words"
    bad = np.flatnonzero(~same)
    if not bad.size:
        return ""
    lost = bad[got[bad, 0] == 0xFFFFFFFF]
    if lost.size:
        return f"{lost.size} rays miss where the checker hits (first {lost[0]})"
    # fp64 Moeller-Trumbore with edges that are eps wide: rays of this campaign run ALONG edges and through vertices, where an
    # fp32 pair test reached through a conservative box may accept what the exact boxes never let it see
    v = verts[:, :3].astype(np.float64)
    tri = v[indices.reshape(-1, 3).astype(np.int64)]
    e1, e2 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    for k in bad[:64]:
        o, d = rays["origin"][k].astype(np.float64), rays["dir"][k].astype(np.float64)
        d = np.where(np.abs(d) < 1e-10, np.copysign(1e-10, d), d)
        p = np.cross(d, e2)
        det = (e1 * p).sum(1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            tv = o - tri[:, 0]
            u = (tv * p).sum(1) * inv
            q = np.cross(tv, e1)
            vv = (q * d).sum(1) * inv
            t = (e2 * q).sum(1) * inv
        # fp32 pair tests of far, small triangles err by ~1e-3 in barycentrics, those of slivers (edge-on, nearly collinear: tiny
        # determinant) by far more: the width grows with the conditioning. Gross errors are what this looks for.
        eps = 2e-3 + 1e-6 * np.linalg.norm(e1, axis=1) * np.linalg.norm(e2, axis=1) / np.maximum(np.abs(det), 1e-300)
        # ... and with distance over size: the hit point is known to ~16 ulp of the largest coordinate or distance involved, which a
        # thin triangle's height turns into a barycentric error
        longest = np.maximum(np.maximum(np.linalg.norm(e1, axis=1), np.linalg.norm(e2, axis=1)), np.linalg.norm(e2 - e1, axis=1))
        height = np.linalg.norm(np.cross(e1, e2), axis=1) / np.maximum(longest, 1e-300)
        reach = max(float(np.abs(o).max()), float(np.abs(v).max()))
        with np.errstate(invalid="ignore"):
            eps = eps + 16 * 1.2e-7 * np.maximum(reach, np.where(np.isfinite(t), np.abs(t), 0.0)) / np.maximum(height, 1e-300)
        degenerate = np.linalg.norm(e1, axis=1) * np.linalg.norm(e2, axis=1) > 1e6 * np.abs(det)  # edge-on / collinear: fp32 cannot decide
        inside = (u >= -eps) & (vv >= -eps) & (u + vv <= 1 + eps) & np.isfinite(t)
        lo_t, hi_t = float(rays["minT"][k]), float(rays["maxT"][k])
        ok_t = inside & (t >= lo_t * (1 - 1e-4) - 1e-30) & (t <= hi_t * (1 + 1e-4))
        g = int(got[k, 0])
        gt = float(got[k, 1:2].copy().view(np.float32)[0])
        if want[k, 0] != 0xFFFFFFFF:
            # both hit, different triangles: north_star's bar is the distance (a tie, or a grazing triangle whose fp32 pair test
            # the exact boxes never reached: its reported t may be off its fp64 t by more than the two answers differ)
            wt = float(want[k, 1:2].copy().view(np.float32)[0])
            if abs(gt - wt) <= 1e-4 * abs(wt):
                continue
            # ... or a NEARER hit the exact boxes let slip (they are not watertight at corners: a ray through a vertex): genuine by fp64
            if gt < wt and g < tri.shape[0] and (ok_t[g] or degenerate[g]):
                continue
            return f"ray {k}: triangle {g} at t {gt}, the checker's triangle {int(want[k, 0])} at t {wt}"
        if g >= tri.shape[0] or not (ok_t[g] or degenerate[g]):
            return f"ray {k}: the checker misses and triangle {g} is not hit even with eps-wide edges"
    return ""


def to_u32(a):
    return np.ascontiguousarray(a).view(np.uint32).reshape(-1, 4)


def trace_device(scene, env, rays):
    if not torch.cuda.is_available():  # dry run over the CPU test build (tests/test_library_on_cpu.py): "device" memory is host memory
        flat = rays.view(np.float32).reshape(-1).copy()
        res = np.full(max(len(rays), 1) * 4, 7.0, np.float32)
        rb.trace_device(scene, env, [(flat.ctypes.data, res.ctypes.data, len(rays))])
        rb.sync()
        return res[: len(rays) * 4].view(np.uint32).reshape(-1, 4)
    d_rays = torch.from_numpy(rays.view(np.float32).reshape(-1).copy()).cuda()
    d_res = torch.full((max(len(rays), 1) * 4,), 7.0, dtype=torch.float32, device="cuda")
    rb.trace_device(scene, env, [(d_rays.data_ptr(), d_res.data_ptr(), len(rays))])
    torch.cuda.synchronize()
    return d_res[: len(rays) * 4].cpu().numpy().view(np.uint32).reshape(-1, 4)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=240.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--rays", type=int, default=20000)
    ap.add_argument("--big", type=float, default=0.02, help="share of rounds with a 20-250 K-triangle soup or mesh")
    ap.add_argument("--devices", type=int, default=1, help="the calling thread's device set: scenes replicated, HOST streams dealt over them")
    args = ap.parse_args()
    global BIG_SHARE
    BIG_SHARE = args.big
    if torch.cuda.is_available():
        torch.cuda.set_device(0)
    rb.init(list(range(args.devices)) if args.devices > 1 else 0)
    rb.set_tuning(**DEFAULT)
    if args.devices > 1:
        rb.set_tuning(host_taper=1)  # with RACC_B200_HOST_CHUNK=2048 in the environment: every device of the set gets chunks of each call
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    rounds = rays_total = builds_compared = frames = quantised = counted = 0
    kinds = {}
    failures = []
    while time.time() - t0 < args.seconds and len(failures) < 5:
        seed = int(rng.integers(0, 2 ** 31))
        r = np.random.default_rng(seed)
        kind, verts, indices = scene_family(r)
        what = f"round {rounds} seed {seed} {kind} {len(indices) // 3} triangles"
        try:
            rb.set_tuning(build_device=int(r.choice([0, 2, 3])))
            scene = rb.create_scene(verts, indices)
            rb.set_tuning(build_device=3)
        except Exception as e:  # noqa: BLE001
            rb.set_tuning(build_device=3)
            failures.append(f"{what}: scene creation failed: {e}")
            rounds += 1
            continue
        env_img = r.random((int(r.integers(1, 9)), int(r.integers(1, 9)), 4)).astype(np.float32) if r.random() < 0.6 else None
        env = rb.create_environment(env_img) if env_img is not None else None
        nodes, pairs, remap = scene.download()
        images = oracle.SceneImages(nodes, pairs, remap, env_img)
        if r.random() < 0.25:  # both builders, same images
            rb.set_tuning(build_device=0)
            other = rb.create_scene(verts, indices)
            rb.set_tuning(build_device=2)
            third = rb.create_scene(verts, indices)
            rb.set_tuning(build_device=3)
            for o in (other, third):
                n2, p2, m2 = o.download()
                if not (n2.tobytes() == nodes.tobytes() and p2.tobytes() == pairs.tobytes() and m2.tobytes() == remap.tobytes()):
                    failures.append(f"{what}: host and device builders disagree")
                o.destroy()
            builds_compared += 1
        rays = rays_for(r, verts, indices, int(r.integers(1, args.rays)))
        want = to_u32(oracle.traverse(images, rays))
        got = trace_device(scene, env, rays)
        if not np.array_equal(got, want):
            bad = np.flatnonzero((got != want).any(1))
            failures.append(f"{what}: default tuning, {bad.size}/{len(rays)} rays differ, first {bad[0]}: got {got[bad[0]]} want {want[bad[0]]} ray {rays[bad[0]]}")
        if torch.cuda.is_available() and r.random() < 0.3:  # several DEVICE streams in one counted launch: results and the visit counters
            cuts = np.sort(r.integers(0, len(rays) + 1, int(r.integers(1, 5))))
            parts = [np.ascontiguousarray(p) for p in np.split(rays, cuts)]
            d_parts = [torch.from_numpy(p.view(np.float32).reshape(-1).copy()).cuda() for p in parts]
            d_outs = [torch.full((max(len(p), 1) * 4,), 7.0, dtype=torch.float32, device="cuda") for p in parts]
            cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
            rb.trace_device(scene, env, [(a.data_ptr(), b.data_ptr(), len(p)) for a, b, p in zip(d_parts, d_outs, parts)], counters_ptr=cnt.data_ptr(), detail=True)
            torch.cuda.synchronize()
            gotm = np.concatenate([b[: len(p) * 4].cpu().numpy().view(np.uint32).reshape(-1, 4) for b, p in zip(d_outs, parts)])
            _, oc = oracle.traverse(images, rays, counters=True)
            c = cnt.cpu().numpy()
            sums = (len(rays), int((want[:, 0] != 0xFFFFFFFF).sum()), int(oc["inner"].astype(np.int64).sum()), int(oc["pairs"].astype(np.int64).sum()),
                    int(oc["pushes"].astype(np.int64).sum()), int(oc["leaves"].astype(np.int64).sum()))
            if not np.array_equal(gotm, want) or tuple(int(x) for x in c[:6]) != sums:
                failures.append(f"{what}: counted launch over DEVICE streams {[len(p) for p in parts]}: results equal {np.array_equal(gotm, want)}, counters {c[:6].tolist()} vs {sums}")
            counted += 1
        shapes = SHAPES if torch.cuda.is_available() else [x for x in SHAPES if x.get("smem_nodes", 0) == 0]  # no TMA on the CPU build
        shape = shapes[int(r.integers(0, len(shapes)))]
        rb.set_tuning(**{**DEFAULT, **shape})
        try:
            got2 = trace_device(scene, env, rays)
        finally:
            rb.set_tuning(**DEFAULT)
        if not np.array_equal(got2, want):
            bad = np.flatnonzero((got2 != want).any(1))
            failures.append(f"{what}: tuning {shape}, {bad.size}/{len(rays)} rays differ, first {bad[0]}")
        if r.random() < 0.5:  # quantised nodes (variant 4, opt-in): north_star's bar instead of bit-exactness
            rb.set_tuning(**{**DEFAULT, "variant": 4})
            try:
                got4 = trace_device(scene, env, rays)
            finally:
                rb.set_tuning(**DEFAULT)
            msg = quantised_check(got4, want, rays, verts, indices)
            if msg:
                failures.append(f"{what}: variant 4: {msg}")
            quantised += 1
        cuts = np.sort(r.integers(0, len(rays) + 1, int(r.integers(0, 6))))
        parts = [np.ascontiguousarray(p) for p in np.split(rays, cuts)]
        outs = [np.zeros(len(p), dtype=rb.RESULT_DTYPE) for p in parts]
        rb.trace_host_ptrs(scene, env, [(p.ctypes.data, o.ctypes.data, len(p)) for p, o in zip(parts, outs)])
        rb.sync()
        got3 = np.concatenate([to_u32(o) for o in outs if len(o)] or [np.zeros((0, 4), np.uint32)])
        if not np.array_equal(got3, want):
            failures.append(f"{what}: HOST streams {[len(p) for p in parts]} differ")
        rays_total += 3 * len(rays)
        if env is not None and r.random() < 0.2 and len(indices) // 3 <= 3000:  # a small frame through both renderer forms
            nt = len(indices) // 3
            tri = verts[indices.reshape(-1, 3).astype(np.int64), :3].astype(np.float64)
            gn = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
            gn /= np.maximum(np.linalg.norm(gn, axis=1, keepdims=True), 1e-300)
            tri_normals = np.zeros((nt, 4), np.float32)
            tri_normals[:, :3] = gn
            normals = np.zeros((verts.shape[0], 4), np.float32)
            normals[:, :3] = r.normal(size=(verts.shape[0], 3))
            normals[:: 53, :3] = 0.0
            materials = np.concatenate([r.uniform(0.05, 0.99, (4, 3)), r.choice([0.4, 1 / 1.5, 1.0, 1.5, 2.4], (4, 1))], 1).astype(np.float32)
            tri_materials = r.integers(0, 6, nt).astype(np.uint16)
            lo, hi = verts[:, :3].min(0), verts[:, :3].max(0)
            w, h = int(r.integers(8, 97)), int(r.integers(8, 65))
            from rayaccel_b200 import scene_io
            cam = scene_io.Camera.look_at((lo + (hi - lo) * r.uniform(-0.3, 1.3, 3)).astype(np.float32), ((lo + hi) / 2).astype(np.float32),
                                          np.array([0.0, 1.0, 0.0], np.float32), float(r.uniform(20, 100)), w, h)
            spp, depth, fseed = int(r.integers(1, 5)), int(r.integers(0, 9)), int(r.integers(0, 100))
            shading = rb.create_shading(normals, tri_normals, tri_materials, materials)
            sh = oracle.Shading(indices, normals, tri_normals, tri_materials, materials)
            want_fb, want_waves = oracle.path_trace(images, sh, cam, w, h, spp, depth, fseed)
            for form in (0, 1):
                rb.set_tuning(path_stream=form)
                try:
                    fb, waves = rb.path_trace(scene, env, shading, cam, w, h, spp, depth, fseed, batch_spp=int(r.integers(0, 3)))
                finally:
                    rb.set_tuning(path_stream=0)
                if waves != [int(x) for x in want_waves] or fb.tobytes() != want_fb.tobytes():
                    failures.append(f"{what}: frame {w}x{h} spp {spp} depth {depth} seed {fseed} form {form} differs from the checker")
            ww, hh = max(8, w // 2), max(8, h // 2)
            want_w, want_ww = oracle.whitted_trace(images, sh, cam, ww, hh, 1, min(depth, 5), fseed)
            fbw, wavesw = rb.whitted_trace(scene, env, shading, cam, ww, hh, 1, min(depth, 5), fseed)
            if wavesw != [int(x) for x in want_ww] or fbw.tobytes() != want_w.tobytes():
                failures.append(f"{what}: Whitted frame {ww}x{hh} depth {min(depth, 5)} seed {fseed} differs from the checker")
            shading.destroy()
            frames += 1
        kinds[kind] = kinds.get(kind, 0) + 1
        if env is not None:
            env.destroy()
        scene.destroy()
        rounds += 1
    print(f"fuzz: {rounds} rounds in {time.time() - t0:.0f} s, {rays_total} rays traced and compared, {builds_compared} builder comparisons, {frames} frames (both path-tracer forms + Whitted), {quantised} rounds also on quantised nodes, {counted} counted multi-stream launches, families { {str(k): v for k, v in kinds.items()} }")
    for f in failures:
        print("FAIL", f)
    print("fuzz: ok" if not failures else f"fuzz: {len(failures)} failures")
    # back to the library's defaults (DEFAULT above pins sort / smem_stack off for the bit-exact comparisons): a test suite may go on
    rb.set_tuning(**{**DEFAULT, "sort": 2, "smem_stack": -1, "build_device": 3, "path_stream": 0, "host_taper": 256})
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
