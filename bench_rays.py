"""Synthetic ray streams of the benchmark workload, generated on the host with numpy.

Neutral input generator: it imports neither the engine (rayaccel_b200/) nor the checker (oracle/). BOTH arms of bench.py
build their batch with it -- the engine arm from the hits the GPU returns, `--impl reference` from the hits the CPU path
returns; the two are bit-identical (that is the parity statement), so both arms trace exactly the same ray bytes, which
bench.py records as a digest in `config`. Camera and bounce models follow the reference's example renderer
(Renderer/Camera.cpp:13-25,55-114; Renderer/PathTracingRenderer.cpp:405-422) with numpy's PCG64 for the jitter and
the hemisphere samples. oracle/raygen.py re-exports these functions for the tests.
"""
from __future__ import annotations

import hashlib
import math

import numpy as np

# racc::Ray (RayAccelerator.h:59-64) and racc::Result (:66-76)
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("minT", "<f4"), ("dir", "<f4", 3), ("maxT", "<f4")])
RESULT_DTYPE = np.dtype([("triangle", "<u4"), ("a", "<f4"), ("b", "<f4"), ("c", "<f4")])
INVALID = 0xFFFFFFFF


def look_at(origin, target, up, fov_degrees, width, height):
    """Camera::lookAt (Camera.cpp:13-25) -> dict(origin, view, right, up) in float32."""
    f32 = np.float32
    origin, target, up = (np.asarray(v, f32) for v in (origin, target, up))
    nrm = lambda v: (v / f32(math.sqrt(float(np.dot(v, v))))).astype(f32)  # noqa: E731
    fwd = nrm(target - origin)
    right = nrm(np.cross(fwd, up).astype(f32))
    cup = np.cross(right, fwd).astype(f32)
    ey = f32(math.tan(0.5 * fov_degrees * (math.pi / 180.0)))
    ex = f32(ey * (f32(width) / f32(height)))
    return dict(origin=origin, view=(fwd + right * ex + cup * ey).astype(f32),
                right=(right * f32(-2.0 / width) * ex).astype(f32), up=(cup * f32(-2.0 / height) * ey).astype(f32))


def primary_rays(cam, width, height, spp=1, seed=0, rows=None):
    """Primary rays (minT 0, maxT 1e6) for the given pixel rows (default: all), sample-major."""
    rows = np.arange(height) if rows is None else np.asarray(rows)
    rng = np.random.default_rng(seed) if seed else None
    out = []
    for _ in range(spp):
        ys, xs = np.meshgrid(rows.astype(np.float32), np.arange(width, dtype=np.float32), indexing="ij")
        if rng is None:
            jx = jy = np.float32(0.5)
        else:
            jx = rng.random(xs.shape, dtype=np.float32)
            jy = rng.random(xs.shape, dtype=np.float32)
        px, py = xs + jx, ys + jy
        d = cam["view"][None, None, :] + cam["up"][None, None, :] * py[..., None] + cam["right"][None, None, :] * px[..., None]
        d = (d / np.linalg.norm(d, axis=2, keepdims=True)).astype(np.float32).reshape(-1, 3)
        r = np.zeros(d.shape[0], dtype=RAY_DTYPE)
        r["origin"], r["dir"], r["minT"], r["maxT"] = cam["origin"], d, 0.0, 1e6
        out.append(r)
    return np.concatenate(out)


def bounce_rays(verts4, indices, rays, results, seed):
    """One diffuse bounce for every hit, in arrival order (PathTracingRenderer.cpp:405-422)."""
    verts = np.asarray(verts4, np.float32).reshape(-1, 4)[:, :3]
    tri = np.asarray(indices, np.uint32).reshape(-1, 3)
    hit = results["triangle"] != INVALID
    r, res = rays[hit], results[hit]
    t = tri[res["triangle"]]
    p0, p1, p2 = verts[t[:, 0]], verts[t[:, 1]], verts[t[:, 2]]
    n = np.cross(p1 - p0, p2 - p0)
    n /= np.maximum(np.linalg.norm(n, axis=1, keepdims=True), 1e-30)
    flip = np.sum(n * r["dir"], axis=1) > 0
    n[flip] = -n[flip]
    hitp = r["origin"] + r["dir"] * res["a"][:, None]
    rng = np.random.default_rng(seed)
    u1, u2 = rng.random(n.shape[0]), rng.random(n.shape[0])
    rr, phi = np.sqrt(u1), 2.0 * np.pi * u2
    lx, ly, lz = rr * np.cos(phi), rr * np.sin(phi), np.sqrt(np.maximum(0.0, 1.0 - u1))
    sgn = np.copysign(1.0, n[:, 2])
    a = -1.0 / (sgn + n[:, 2])
    b = n[:, 0] * n[:, 1] * a
    t1 = np.stack([1.0 + sgn * n[:, 0] ** 2 * a, sgn * b, -sgn * n[:, 0]], axis=1)
    t2 = np.stack([b, sgn + n[:, 1] ** 2 * a, -n[:, 1]], axis=1)
    d = lx[:, None] * t1 + ly[:, None] * t2 + lz[:, None] * n
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    out = np.zeros(n.shape[0], dtype=RAY_DTYPE)
    out["origin"] = (hitp + 1e-4 * n).astype(np.float32)
    out["dir"] = d.astype(np.float32)
    out["minT"], out["maxT"] = 1e-3, 1e6
    return out


def wavefront(scene_file, width, height, spp, bounces, seed, trace, rows=None):
    """The benchmark batch: primary rays of width x height x spp (jitter seed `seed`) followed by `bounces` diffuse bounces,
    bounce b built from the hits `trace(rays) -> RESULT_DTYPE array` returns for the previous wave (seed + 1 + b).
    Returns the list of ray arrays and the list of result arrays (the last wave's results included)."""
    sf = scene_file
    cam = look_at(sf.cam_origin, sf.cam_target, sf.cam_up, sf.cam_fov, width, height)
    streams = [primary_rays(cam, width, height, spp, seed, rows=rows)]
    results = []
    for b in range(bounces + 1):
        results.append(trace(streams[-1]))
        if b == bounces:
            break
        streams.append(bounce_rays(sf.vertices, sf.indices, streams[-1], results[-1], seed + 1 + b))
    return streams, results


def digest(streams) -> str:
    """Short hash of the ray bytes of a batch: equal on both arms of bench.py when they traced the same rays."""
    h = hashlib.sha256()
    for s in streams:
        h.update(np.ascontiguousarray(s).view(np.uint8).reshape(-1).data)
    return h.hexdigest()[:16]
