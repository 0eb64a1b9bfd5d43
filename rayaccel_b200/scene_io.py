"""Loader for the reference's `battlefield.bin` scene container and its camera model.

File layout: /root/reference/Renderer/main.cpp:117-191 (60-byte header, then indices, per-triangle
materials and normals, vertices, vertex normals, texcoords, RGBA32F light probe). Camera:
/root/reference/Renderer/Camera.cpp:13-25. data/battlefield.bin is the reference's own fixture,
copied unchanged (it is data, not source).
"""
from __future__ import annotations

import math
import os
import struct
from dataclasses import dataclass

import numpy as np

DEFAULT_SCENE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data", "battlefield.bin")


@dataclass
class SceneFile:
    max_depth: int
    viewport: tuple[int, int]
    vertices: np.ndarray        # (V, 4) float32, w = 1
    indices: np.ndarray         # (3T,) uint32
    materials: np.ndarray       # (T,) uint16
    triangle_normals: np.ndarray  # (T, 4) float32
    normals: np.ndarray         # (V, 4) float32
    texcoords: np.ndarray       # (V, 2) float32
    environment: np.ndarray     # (H, W, 4) float32
    cam_origin: np.ndarray
    cam_target: np.ndarray
    cam_up: np.ndarray
    cam_fov: float

    @property
    def triangle_count(self) -> int:
        return self.indices.shape[0] // 3


def load_scene(path: str = DEFAULT_SCENE) -> SceneFile:
    raw = open(path, "rb").read()
    max_depth, nv, nt, vw, vh, ew, eh = struct.unpack_from("<IIIHHHH", raw, 0)
    cam = struct.unpack_from("<10f", raw, 20)
    off = 60

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=off).copy()
        off += a.nbytes
        return a

    indices = take(np.uint32, nt * 3)
    materials = take(np.uint16, nt)
    tri_normals = take(np.float32, nt * 4).reshape(nt, 4)
    vertices = take(np.float32, nv * 4).reshape(nv, 4)
    normals = take(np.float32, nv * 4).reshape(nv, 4)
    texcoords = take(np.float32, nv * 2).reshape(nv, 2)
    env = take(np.float32, ew * eh * 4).reshape(eh, ew, 4)
    if off != len(raw):
        raise ValueError(f"{path}: {len(raw) - off} trailing bytes, not a battlefield.bin container")
    return SceneFile(max_depth, (vw, vh), vertices, indices, materials, tri_normals, normals, texcoords, env,
                     np.array(cam[0:3], np.float32), np.array(cam[3:6], np.float32), np.array(cam[6:9], np.float32), float(cam[9]))


@dataclass
class Camera:
    """Camera::lookAt of the reference (Camera.cpp:13-25), evaluated in float32."""
    origin: np.ndarray
    view: np.ndarray
    right: np.ndarray
    up: np.ndarray

    @staticmethod
    def look_at(origin, target, up, fov_degrees: float, width: int, height: int) -> "Camera":
        f32 = np.float32
        origin = np.asarray(origin, f32)
        target = np.asarray(target, f32)
        up = np.asarray(up, f32)

        def normalize(v):
            return (v / f32(math.sqrt(float(np.dot(v, v))))).astype(f32)

        forward = normalize(target - origin)
        right = normalize(np.cross(forward, up).astype(f32))
        cam_up = np.cross(right, forward).astype(f32)
        aspect = f32(width) / f32(height)
        ext_y = f32(math.tan(0.5 * fov_degrees * (math.pi / 180.0)))
        ext_x = f32(ext_y * aspect)
        return Camera(origin,
                      (forward + right * ext_x + cam_up * ext_y).astype(f32),
                      (right * f32(-2.0 / width) * ext_x).astype(f32),
                      (cam_up * f32(-2.0 / height) * ext_y).astype(f32))

    @staticmethod
    def for_scene(scene: SceneFile, width: int, height: int) -> "Camera":
        return Camera.look_at(scene.cam_origin, scene.cam_target, scene.cam_up, scene.cam_fov, width, height)


def synthetic_triangles(n_triangles: int, seed: int = 7, extent: float = 1000.0, edge: float = 2.0):
    """BASELINE.json config 5: random triangle soup. Returns (verts4, indices)."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(0.0, extent, size=(n_triangles, 1, 3)).astype(np.float32)
    offs = (rng.uniform(-1.0, 1.0, size=(n_triangles, 3, 3)) * edge).astype(np.float32)
    verts = np.ones((n_triangles * 3, 4), dtype=np.float32)
    verts[:, :3] = (centres + offs).reshape(-1, 3)
    indices = np.arange(n_triangles * 3, dtype=np.uint32)
    return verts, indices
