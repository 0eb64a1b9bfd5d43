"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the parity checker.

`oracle/liboracle.so` is our CPU restatement of the reference hot path (racc_oracle.c, which
cites /root/reference/RayAccelerator/Kernels.h line by line) and `oracle/_ref/libracc_ref.so`
is the UNMODIFIED reference scene builder + light-probe sampler compiled with stand-in headers
(oracle/ref_shim/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package; the product (rayaccel_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libracc_ref.so")

INVALID = 0xFFFFFFFF

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("minT", "<f4"), ("dir", "<f4", 3), ("maxT", "<f4")])
RESULT_DTYPE = np.dtype([("triangle", "<u4"), ("a", "<f4"), ("b", "<f4"), ("c", "<f4")])
COUNTER_DTYPE = np.dtype([("inner", "<u2"), ("pairs", "<u2"), ("max_stack", "<u2"), ("hit", "<u2"), ("pushes", "<u2"), ("leaves", "<u2")])


def build(ref: bool = True) -> None:
    """Compile the checker (and, when /root/reference is present, oracle/_ref)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"] + (["ref", "kernel", "renderer"] if ref else []), stdout=subprocess.DEVNULL)


class _Scene(ctypes.Structure):
    _fields_ = [
        ("nodes", ctypes.c_void_p), ("node_count", ctypes.c_uint32),
        ("pairs", ctypes.c_void_p), ("pair_count", ctypes.c_uint32),
        ("remap", ctypes.c_void_p), ("remap_count", ctypes.c_uint32),
        ("env", ctypes.c_void_p), ("env_width", ctypes.c_uint32), ("env_height", ctypes.c_uint32),
    ]


class _RefImages(ctypes.Structure):
    _fields_ = [
        ("nodes", ctypes.c_void_p), ("nodes_bytes", ctypes.c_uint64),
        ("pairs", ctypes.c_void_p), ("pairs_bytes", ctypes.c_uint64),
        ("remap", ctypes.c_void_p), ("remap_bytes", ctypes.c_uint64),
    ]


_lib = None
_ref = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        _lib = ctypes.CDLL(ORACLE_SO)
        _lib.oracle_acosf.restype = ctypes.c_float
        _lib.oracle_acosf.argtypes = [ctypes.c_float]
    return _lib


def have_ref() -> bool:
    return os.path.exists(REF_SO)


KERNEL_SO = os.path.join(_HERE, "_ref", "libkernel_ref.so")
_kernel = None


def have_ref_kernel() -> bool:
    return os.path.exists(KERNEL_SO)


KERNEL_RELAXED_SO = os.path.join(_HERE, "_ref", "libkernel_ref_relaxed.so")
_kernel_relaxed = None


def ref_kernel_traverse_relaxed(scene: "SceneImages", rays: np.ndarray, threads: int = 0) -> np.ndarray:
    """The reference kernel source under the OTHER built-in model of oracle/ref_shim/opencl_c.h (RACC_SHIM_RELAXED: unfused
    mad / dot, estimate-based native_recip / native_rsqrt): what a different, equally conforming OpenCL implementation
    could return. Only for measuring the distance to the pinned results."""
    global _kernel_relaxed
    if _kernel_relaxed is None:
        lib()
        _kernel_relaxed = ctypes.CDLL(KERNEL_RELAXED_SO)
    return _run_ref_kernel(_kernel_relaxed, scene, rays, threads)


def ref_kernel_traverse(scene: "SceneImages", rays: np.ndarray, threads: int = 1) -> np.ndarray:
    """The reference's OpenCL `traversal` kernel, compiled from its own source text (Kernels.h) on top of
    oracle/ref_shim/opencl_c.h, run on the CPU one work-item per ray; threads > 1 hands batches of 1024
    work-items to that many threads, threads <= 0 uses every online core."""
    global _kernel
    if _kernel is None:
        lib()  # liboracle.so provides oracle_acosf
        _kernel = ctypes.CDLL(KERNEL_SO)
    return _run_ref_kernel(_kernel, scene, rays, threads)


def _run_ref_kernel(_kernel, scene: "SceneImages", rays: np.ndarray, threads: int) -> np.ndarray:
    rays = np.ascontiguousarray(rays)
    assert rays.dtype == RAY_DTYPE
    n = rays.shape[0]
    res = np.zeros(n, dtype=RESULT_DTYPE)
    s = scene.c_struct()
    if threads <= 0:
        threads = os.cpu_count() or 1
    _kernel.ref_kernel_traverse_mt(ctypes.c_void_p(s.nodes), ctypes.c_void_p(s.pairs), ctypes.c_void_p(s.remap), ctypes.c_void_p(s.env),
                                   ctypes.c_uint32(s.env_width), ctypes.c_uint32(s.env_height), _p(rays), ctypes.c_uint32(n), _p(res),
                                   ctypes.c_int(threads))
    return res


def ref() -> ctypes.CDLL:
    global _ref
    if _ref is None:
        _ref = ctypes.CDLL(REF_SO)
    return _ref


def _p(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


class SceneImages:
    """Host copies of the three scene images (+ optional light probe) the oracle traverses."""

    def __init__(self, nodes: np.ndarray, pairs: np.ndarray, remap: np.ndarray, env: np.ndarray | None = None):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1, 16)
        self.pairs = np.ascontiguousarray(pairs, dtype=np.float32).reshape(-1, 12)
        self.remap = np.ascontiguousarray(remap, dtype=np.uint32).reshape(-1)
        self.env = None if env is None else np.ascontiguousarray(env, dtype=np.float32)
        if self.env is not None:
            assert self.env.ndim == 3 and self.env.shape[2] == 4

    def c_struct(self) -> _Scene:
        s = _Scene()
        s.nodes, s.node_count = _p(self.nodes), self.nodes.shape[0]
        s.pairs, s.pair_count = _p(self.pairs), self.pairs.shape[0]
        s.remap, s.remap_count = _p(self.remap), self.remap.shape[0]
        if self.env is not None:
            s.env, s.env_height, s.env_width = _p(self.env), self.env.shape[0], self.env.shape[1]
        return s

    def digest(self) -> dict:
        d = (ctypes.c_uint64 * 16)()
        s = self.c_struct()
        rc = lib().oracle_scene_digest(ctypes.byref(s), d)
        if rc:
            raise RuntimeError("oracle_scene_digest: malformed scene image")
        return {
            "inner": int(d[0]), "leaves": int(d[1]), "pairs": int(d[2]), "depth": int(d[3]),
            "singletons": int(d[4]), "leaf_hist": [int(d[5 + i]) for i in range(7)], "hash": int(d[12]),
        }


def traverse(scene: SceneImages, rays: np.ndarray, counters: bool = False, threads: int = 0):
    rays = np.ascontiguousarray(rays)
    assert rays.dtype == RAY_DTYPE
    n = rays.shape[0]
    res = np.zeros(n, dtype=RESULT_DTYPE)
    cnt = np.zeros(n, dtype=COUNTER_DTYPE) if counters else None
    s = scene.c_struct()
    rc = lib().oracle_traverse(ctypes.byref(s), _p(rays), ctypes.c_uint32(n), _p(res),
                               _p(cnt) if counters else None, ctypes.c_int(threads))
    if rc:
        raise RuntimeError("oracle_traverse failed (stack overflow or malformed scene)")
    return (res, cnt) if counters else res


def traverse_avx2(scene: SceneImages, rays: np.ndarray, threads: int = 0) -> np.ndarray:
    """CPU BASELINE (bench.py only): the same traversal with an AVX2 node test, all host threads by default."""
    rays = np.ascontiguousarray(rays)
    assert rays.dtype == RAY_DTYPE
    n = rays.shape[0]
    res = np.zeros(n, dtype=RESULT_DTYPE)
    s = scene.c_struct()
    rc = lib().oracle_traverse_avx2(ctypes.byref(s), _p(rays), ctypes.c_uint32(n), _p(res), ctypes.c_int(threads))
    if rc:
        raise RuntimeError("oracle_traverse_avx2 failed (stack overflow or malformed scene)")
    return res


def env_sample(env: np.ndarray, dirs: np.ndarray) -> np.ndarray:
    env = np.ascontiguousarray(env, dtype=np.float32)
    dirs = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
    out = np.zeros((dirs.shape[0], 3), dtype=np.float32)
    f = lib().oracle_env_sample
    for i in range(dirs.shape[0]):
        f(_p(env), ctypes.c_uint32(env.shape[1]), ctypes.c_uint32(env.shape[0]), _p(dirs[i]), _p(out[i]))
    return out


def acosf(x: np.ndarray) -> np.ndarray:
    f = lib().oracle_acosf
    return np.array([f(float(v)) for v in np.asarray(x, dtype=np.float32).ravel()], dtype=np.float32)


def brute_f64(verts4: np.ndarray, indices: np.ndarray, rays: np.ndarray, threads: int = 0):
    verts4 = np.ascontiguousarray(verts4, dtype=np.float32).reshape(-1, 4)
    indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
    rays = np.ascontiguousarray(rays)
    n = rays.shape[0]
    t = np.zeros(n, dtype=np.float64)
    tri = np.zeros(n, dtype=np.uint32)
    lib().oracle_brute_f64(_p(verts4), ctypes.c_uint32(verts4.shape[0]), _p(indices), ctypes.c_uint32(indices.shape[0] // 3),
                           _p(rays), ctypes.c_uint32(n), _p(t), _p(tri), ctypes.c_int(threads))
    return t, tri


def tri_t_f64(verts4: np.ndarray, indices: np.ndarray, rays: np.ndarray, tri_ids: np.ndarray) -> np.ndarray:
    verts4 = np.ascontiguousarray(verts4, dtype=np.float32).reshape(-1, 4)
    indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
    rays = np.ascontiguousarray(rays)
    tri_ids = np.ascontiguousarray(tri_ids, dtype=np.uint32)
    t = np.zeros(rays.shape[0], dtype=np.float64)
    lib().oracle_tri_t_f64(_p(verts4), _p(indices), _p(rays), _p(tri_ids), ctypes.c_uint32(rays.shape[0]), _p(t))
    return t


# ---- the unmodified reference (oracle/_ref) ------------------------------------------------

def ref_cpu_query(verts4: np.ndarray, indices: np.ndarray, env: np.ndarray | None, rays: np.ndarray, threads: int = 0) -> np.ndarray:
    """The reference's CPU query path, executeRayQueryCPU (Scene.cpp:374-484), run from its own source over the stand-in
    for its binary-only Embree 2.7 (oracle/ref_shim/mini_embree.cpp): RESULT_DTYPE array, Embree conventions (primID =
    original triangle index, u / v = weights of vertices 1 and 2, misses carry the light probe's radiance)."""
    verts4 = np.ascontiguousarray(verts4, dtype=np.float32).reshape(-1, 4)
    indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
    rays = np.ascontiguousarray(rays)
    assert rays.dtype == RAY_DTYPE
    res = np.zeros(rays.shape[0], dtype=RESULT_DTYPE)
    e = None if env is None else np.ascontiguousarray(env, dtype=np.float32)
    fn = ref().ref_cpu_query
    fn.restype = ctypes.c_int
    rc = fn(_p(verts4), ctypes.c_uint32(verts4.shape[0]), _p(indices), ctypes.c_uint32(indices.shape[0]),
            _p(e) if e is not None else None, ctypes.c_uint32(e.shape[1] if e is not None else 0),
            ctypes.c_uint32(e.shape[0] if e is not None else 0), _p(rays), ctypes.c_uint32(rays.shape[0]), _p(res), ctypes.c_int(threads))
    if rc:
        raise RuntimeError("ref_cpu_query failed")
    return res


def have_ref_cpu_query() -> bool:
    return have_ref() and hasattr(ref(), "ref_cpu_query")


def ref_build_scene(verts4: np.ndarray, indices: np.ndarray) -> SceneImages:
    """racc::createScene() of the unmodified reference -> its GPU upload images."""
    verts4 = np.ascontiguousarray(verts4, dtype=np.float32).reshape(-1, 4)
    indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
    img = _RefImages()
    rc = ref().ref_build_scene(_p(verts4), ctypes.c_uint32(verts4.shape[0]), _p(indices),
                               ctypes.c_uint32(indices.shape[0]), ctypes.byref(img))
    if rc:
        raise RuntimeError("ref_build_scene failed")
    try:
        def grab(ptr, nbytes, dtype):
            buf = (ctypes.c_char * nbytes).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype).copy()
        return SceneImages(grab(img.nodes, img.nodes_bytes, np.float32),
                           grab(img.pairs, img.pairs_bytes, np.float32),
                           grab(img.remap, img.remap_bytes, np.uint32))
    finally:
        ref().ref_free_scene_images(ctypes.byref(img))


def ref_env_sample(env: np.ndarray, dirs: np.ndarray) -> np.ndarray:
    """racc_internal::sample() of the unmodified reference (Environment.h:27-82)."""
    env = np.ascontiguousarray(env, dtype=np.float32)
    d4 = np.zeros((np.asarray(dirs).reshape(-1, 3).shape[0], 4), dtype=np.float32)
    d4[:, :3] = np.asarray(dirs, dtype=np.float32).reshape(-1, 3)
    out = np.zeros_like(d4)
    rc = ref().ref_env_sample(_p(env), ctypes.c_uint32(env.shape[1]), ctypes.c_uint32(env.shape[0]),
                              _p(d4), ctypes.c_uint32(d4.shape[0]), _p(out))
    if rc:
        raise RuntimeError("ref_env_sample failed")
    return out[:, :3]


# ---- wavefront path tracer: checker for rayaccel_b200/csrc/pathtrace.cu -----------------

class _Shading(ctypes.Structure):
    _fields_ = [
        ("indices", ctypes.c_void_p), ("triangle_count", ctypes.c_uint32),
        ("normals4", ctypes.c_void_p), ("triangle_normals4", ctypes.c_void_p),
        ("triangle_materials", ctypes.c_void_p), ("materials_ke4", ctypes.c_void_p), ("material_count", ctypes.c_uint32),
    ]


class _Camera(ctypes.Structure):
    _fields_ = [("origin", ctypes.c_float * 3), ("view", ctypes.c_float * 3), ("right", ctypes.c_float * 3), ("up", ctypes.c_float * 3)]


# the four materials the reference assigns to battlefield.bin (Renderer/main.cpp:165-168): {r, g, b, eta}
BATTLEFIELD_MATERIALS = np.array([[0.8, 0.8, 0.8, 1.0 / 1.4], [0.1, 0.1, 0.1, 1.0 / 1.4], [0.6, 0.6, 0.6, 1.0 / 1.2], [0.3, 0.3, 0.3, 1.0 / 1.2]],
                                 dtype=np.float32)


class Shading:
    """Host copies of what the reference's example path tracer shades with (Renderer/SceneData.h)."""

    def __init__(self, indices, normals4, triangle_normals4, triangle_materials, materials_ke4=BATTLEFIELD_MATERIALS):
        self.indices = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        self.normals = np.ascontiguousarray(normals4, dtype=np.float32).reshape(-1, 4)
        self.triangle_normals = np.ascontiguousarray(triangle_normals4, dtype=np.float32).reshape(-1, 4)
        self.triangle_materials = np.ascontiguousarray(triangle_materials, dtype=np.uint16).reshape(-1)
        self.materials = np.ascontiguousarray(materials_ke4, dtype=np.float32).reshape(-1, 4)
        assert self.indices.shape[0] == 3 * self.triangle_normals.shape[0] == 3 * self.triangle_materials.shape[0]

    def c_struct(self) -> _Shading:
        return _Shading(self.indices.ctypes.data, self.triangle_materials.shape[0], self.normals.ctypes.data,
                        self.triangle_normals.ctypes.data, self.triangle_materials.ctypes.data, self.materials.ctypes.data,
                        self.materials.shape[0])


def _camera_struct(cam) -> _Camera:
    get = (lambda k: cam[k]) if isinstance(cam, dict) else (lambda k: getattr(cam, k))
    c = _Camera()
    for k in range(3):
        c.origin[k], c.view[k], c.right[k], c.up[k] = (float(get(n)[k]) for n in ("origin", "view", "right", "up"))
    return c


def material_sample(ke, rnd, normal, wo):
    """oracle_material_sample for n lanes: (n,3) rnd / normal / wo -> (wi, color)."""
    rnd, normal, wo = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (rnd, normal, wo))
    ke = np.ascontiguousarray(ke, dtype=np.float32).reshape(4)
    wi, color = np.zeros_like(rnd), np.zeros_like(rnd)
    fn = lib().oracle_material_sample
    fn.restype = None
    for i in range(rnd.shape[0]):
        fn(_p(ke), _p(rnd[i]), _p(normal[i]), _p(wo[i]), ctypes.c_void_p(wi[i].ctypes.data), ctypes.c_void_p(color[i].ctypes.data))
    return wi, color


def path_trace(scene: SceneImages, shading: Shading, camera, width: int, height: int, spp: int, max_depth: int, seed: int,
               sample_base: int = 0, framebuffer: np.ndarray | None = None, threads: int = 0):
    """oracle_path_trace: returns (framebuffer (H, W, 4) float32 with the radiance sums added, rays traced per depth)."""
    fb = np.zeros((height, width, 4), dtype=np.float32) if framebuffer is None else framebuffer
    assert fb.dtype == np.float32 and fb.shape == (height, width, 4) and fb.flags.c_contiguous
    waves = np.zeros(max_depth + 1, dtype=np.uint64)
    s, sh, cam = scene.c_struct(), shading.c_struct(), _camera_struct(camera)
    rc = lib().oracle_path_trace(ctypes.byref(s), ctypes.byref(sh), ctypes.byref(cam), ctypes.c_uint32(width), ctypes.c_uint32(height),
                                 ctypes.c_uint32(sample_base), ctypes.c_uint32(spp), ctypes.c_uint32(max_depth), ctypes.c_uint32(seed),
                                 _p(fb), _p(waves), ctypes.c_int(threads))
    if rc:
        raise RuntimeError("oracle_path_trace failed")
    return fb, waves


def whitted_trace(scene: SceneImages, shading: Shading, camera, width: int, height: int, spp: int, max_depth: int, seed: int,
                  sample_base: int = 0, framebuffer: np.ndarray | None = None, threads: int = 0):
    """oracle_whitted_trace: returns (framebuffer (H, W, 4) float32 with the radiance sums added, rays traced per depth)."""
    fb = np.zeros((height, width, 4), dtype=np.float32) if framebuffer is None else framebuffer
    assert fb.dtype == np.float32 and fb.shape == (height, width, 4) and fb.flags.c_contiguous
    waves = np.zeros(max_depth + 1, dtype=np.uint64)
    s, sh, cam = scene.c_struct(), shading.c_struct(), _camera_struct(camera)
    rc = lib().oracle_whitted_trace(ctypes.byref(s), ctypes.byref(sh), ctypes.byref(cam), ctypes.c_uint32(width), ctypes.c_uint32(height),
                                    ctypes.c_uint32(sample_base), ctypes.c_uint32(spp), ctypes.c_uint32(max_depth), ctypes.c_uint32(seed),
                                    _p(fb), _p(waves), ctypes.c_int(threads))
    if rc:
        raise RuntimeError("oracle_whitted_trace failed")
    return fb, waves


SHADE_SO = os.path.join(_HERE, "_ref", "libshade_ref.so")
_shade = None


def have_ref_shade() -> bool:
    return os.path.exists(SHADE_SO)


def ref_material_sample(ke, rnd, normal, wo):
    """The reference's UNMODIFIED ReflectiveDiffuseMaterial::sample8 (Renderer/Materials.cpp, compiled into
    oracle/_ref/libshade_ref.so by `make -C oracle renderer`), 8 lanes at a time."""
    global _shade
    if _shade is None:
        _shade = ctypes.CDLL(SHADE_SO)
    rnd, normal, wo = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in (rnd, normal, wo))
    ke = np.ascontiguousarray(ke, dtype=np.float32).reshape(4)
    n = rnd.shape[0]
    wi, color = np.zeros_like(rnd), np.zeros_like(rnd)
    _shade.ref_material_sample(_p(ke), _p(rnd), _p(normal), _p(wo), ctypes.c_uint32(n), _p(wi), _p(color))
    return wi, color


def _shade_lib() -> ctypes.CDLL:
    global _shade
    if _shade is None:
        _shade = ctypes.CDLL(SHADE_SO)
    return _shade


def ref_camera_look_at(origin, target, up, fov: float, width: int, height: int) -> dict:
    """The reference's Camera::lookAt (Renderer/Camera.cpp:13-25), executed: dict(origin, view, right, up)."""
    o, t, u = (np.ascontiguousarray(a, dtype=np.float32).reshape(3) for a in (origin, target, up))
    out = np.zeros(12, dtype=np.float32)
    _shade_lib().ref_camera_look_at(_p(o), _p(t), _p(u), ctypes.c_float(fov), ctypes.c_int(width), ctypes.c_int(height), _p(out))
    return dict(origin=out[0:3].copy(), view=out[3:6].copy(), right=out[6:9].copy(), up=out[9:12].copy())


def ref_generate_tile(camera, tile_x: int, tile_y: int, tile_size: int, viewport_width: int):
    """The reference's generateTileRays + generateTileLightPaths (Camera.cpp:55-114, LightPath.cpp:11-39) for one tile:
    (rays as RAY_DTYPE, light paths as (n, 4) uint32 words: weight rgb bits, pixel)."""
    c = _camera_struct(camera)
    cam = np.array(list(c.origin) + list(c.view) + list(c.right) + list(c.up), dtype=np.float32)
    n = tile_size * tile_size
    rays = np.zeros(n, dtype=RAY_DTYPE)
    paths = np.zeros((n, 4), dtype=np.uint32)
    _shade_lib().ref_generate_tile(_p(cam), ctypes.c_uint(tile_x), ctypes.c_uint(tile_y), ctypes.c_uint(tile_size), ctypes.c_uint(viewport_width),
                                   _p(rays), _p(paths))
    return rays, paths
