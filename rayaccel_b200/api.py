"""Python mirror of the reference's public interface for the intersection path.

Names and argument meaning follow /root/reference/RayAccelerator/RayAccelerator.h:95-115
(`create_scene` = racc::createScene, `create_environment` = racc::createEnvironment,
`destroy`, and ray streams of racc::Ray / racc::Result records). Everything is executed by the
CUDA engine behind the C-ABI in include/racc_b200.h; this file only marshals pointers.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import Counters, EngineError, SceneInfo, StreamDesc
from .scene_io import Camera

INVALID_TRIANGLE = 0xFFFFFFFF  # racc::invalidTriangle, RayAccelerator.h:26

# racc::Ray (RayAccelerator.h:59-64) and racc::Result (:66-76)
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("minT", "<f4"), ("dir", "<f4", 3), ("maxT", "<f4")])
RESULT_DTYPE = np.dtype([("triangle", "<u4"), ("a", "<f4"), ("b", "<f4"), ("c", "<f4")])

NODE_BYTES, PAIR_BYTES = 64, 48


def init(device=None) -> None:
    """racc::init(): sets the calling thread's device set -- one CUDA ordinal, a list of them (the first is the device the
    thread is bound to; scenes are replicated on all, HOST streams dealt over all), or None for the current device."""
    lib = _lib.load()
    if device is None:
        _lib.check(lib.racc_cuda_init(None, 0), "racc_cuda_init")
    else:
        devices = [int(device)] if isinstance(device, (int, np.integer)) else [int(d) for d in device]
        arr = (ctypes.c_int * len(devices))(*devices)
        _lib.check(lib.racc_cuda_init(arr, len(devices)), "racc_cuda_init")


def current_devices() -> list[int]:
    arr = (ctypes.c_int * 64)()
    n = _lib.load().racc_cuda_current_devices(arr, 64)
    if n < 0:
        raise EngineError(_lib.last_error())
    return [int(arr[k]) for k in range(n)]


def thread_release() -> None:
    """Frees the calling thread's staging pipelines and renderer scratch (racc_cuda_thread_release)."""
    _lib.load().racc_cuda_thread_release()


def frame_reduce(stream=None, wait: bool = True):
    """The per-frame hit reduction (racc_cuda_frame_reduce): sums and zeroes the frame records of the thread's device set and,
    after comm_init_rank, of all ranks. wait=True returns {rays, hits, inner_nodes, pairs_tested}; wait=False only enqueues."""
    lib = _lib.load()
    if not wait:
        _lib.check(lib.racc_cuda_frame_reduce(None, _cuda_stream_handle(stream)), "racc_cuda_frame_reduce")
        return None
    c = Counters()
    _lib.check(lib.racc_cuda_frame_reduce(ctypes.byref(c), _cuda_stream_handle(stream)), "racc_cuda_frame_reduce")
    return {"rays": int(c.rays), "hits": int(c.hits), "inner_nodes": int(c.inner_nodes), "pairs_tested": int(c.pairs_tested)}


def comm_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    _lib.check(_lib.load().racc_cuda_comm_unique_id(buf), "racc_cuda_comm_unique_id")
    return buf.raw


def comm_init_rank(unique_id: bytes, rank: int, nranks: int) -> None:
    buf = ctypes.create_string_buffer(bytes(unique_id), 128)
    _lib.check(_lib.load().racc_cuda_comm_init_rank(buf, rank, nranks), "racc_cuda_comm_init_rank")


def comm_destroy() -> None:
    _lib.load().racc_cuda_comm_destroy()


def comm_ranks() -> tuple[int, int]:
    """(rank, number of ranks) of the communicator this process joined; (0, 1) without one."""
    r = ctypes.c_int(0)
    n = _lib.load().racc_cuda_comm_ranks(ctypes.byref(r))
    return int(r.value), int(n)


def gather_results(results_ptr: int, rays_per_rank: int, all_results_ptr: int, stream=None) -> None:
    """racc_cuda_gather_results: every rank's Result slice (rays_per_rank x 16 B, device memory) gathered in rank order into
    all_results_ptr on every rank (ncclAllGather inside the engine library)."""
    _lib.check(_lib.load().racc_cuda_gather_results(ctypes.c_void_p(results_ptr), rays_per_rank, ctypes.c_void_p(all_results_ptr),
                                                    _cuda_stream_handle(stream)), "racc_cuda_gather_results")


def device_count() -> int:
    return _lib.load().racc_cuda_device_count()


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


def _info_dict(info: SceneInfo) -> dict:
    return {
        "node_count": info.node_count, "pair_count": info.pair_count, "real_pair_count": info.real_pair_count,
        "remap_count": info.remap_count, "depth": info.depth, "triangle_count": info.triangle_count,
        "bounds_min": tuple(info.bounds_min), "bounds_max": tuple(info.bounds_max),
    }


class HostImages:
    """The three scene images in host memory (no CUDA involved): racc_cuda_build_images."""

    def __init__(self, vertices: np.ndarray, indices: np.ndarray):
        lib = _lib.load()
        v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 4)
        i = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        self._h = _lib.check_ptr(lib.racc_cuda_build_images(_ptr(v), v.shape[0], _ptr(i), i.shape[0]), "racc_cuda_build_images")
        info = SceneInfo()
        _lib.check(lib.racc_cuda_host_images_get_info(self._h, ctypes.byref(info)), "racc_cuda_host_images_get_info")
        self.info = _info_dict(info)
        self.nodes = np.zeros((info.node_count, 16), dtype=np.float32)
        self.pairs = np.zeros((info.pair_count, 12), dtype=np.float32)
        self.remap = np.zeros(info.remap_count, dtype=np.uint32)
        _lib.check(lib.racc_cuda_host_images_copy(self._h, _ptr(self.nodes), _ptr(self.pairs), _ptr(self.remap)), "racc_cuda_host_images_copy")
        lib.racc_cuda_host_images_destroy(self._h)
        self._h = None


class Scene:
    """racc::Scene. Immutable after creation; may be shared by successive traces."""

    def __init__(self, handle):
        self._h = handle
        info = SceneInfo()
        _lib.check(_lib.load().racc_cuda_scene_get_info(self._h, ctypes.byref(info)), "racc_cuda_scene_get_info")
        self.info = _info_dict(info)

    def download(self):
        """(nodes, pairs, remap) as the kernel sees them (read back from the device)."""
        nodes = np.zeros((self.info["node_count"], 16), dtype=np.float32)
        pairs = np.zeros((self.info["pair_count"], 12), dtype=np.float32)
        remap = np.zeros(self.info["remap_count"], dtype=np.uint32)
        _lib.check(_lib.load().racc_cuda_scene_download(self._h, _ptr(nodes), _ptr(pairs), _ptr(remap)), "racc_cuda_scene_download")
        return nodes, pairs, remap

    def destroy(self) -> None:
        if self._h:
            _lib.load().racc_cuda_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Environment:
    """racc::Environment: an angular-map light probe (RGBA32F)."""

    def __init__(self, handle, width: int, height: int):
        self._h, self.width, self.height = handle, width, height

    def destroy(self) -> None:
        if self._h:
            _lib.load().racc_cuda_env_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def create_scene(vertices: np.ndarray, indices: np.ndarray) -> Scene:
    """racc::createScene(context, vertices, vertexCount, indices, indexCount) (RayAccelerator.h:107).
    vertices: (V,4) float32, indices: (3T,) uint32. Arrays are copied; the caller may free them."""
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    if v.ndim != 2 or v.shape[1] != 4:
        raise ValueError("vertices must be (V, 4) float32 (racc::Vertex)")
    i = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
    if i.shape[0] % 3:
        raise ValueError("indexCount % 3 != 0")  # Scene.cpp:186
    lib = _lib.load()
    return Scene(_lib.check_ptr(lib.racc_cuda_scene_create(_ptr(v), v.shape[0], _ptr(i), i.shape[0]), "racc_cuda_scene_create"))


def create_scene_from_images(nodes: np.ndarray, pairs: np.ndarray, remap: np.ndarray) -> Scene:
    n = np.ascontiguousarray(nodes, dtype=np.float32).reshape(-1, 16)
    p = np.ascontiguousarray(pairs, dtype=np.float32).reshape(-1, 12)
    r = np.ascontiguousarray(remap, dtype=np.uint32).reshape(-1)
    lib = _lib.load()
    return Scene(_lib.check_ptr(lib.racc_cuda_scene_create_from_images(_ptr(n), n.shape[0], _ptr(p), p.shape[0], _ptr(r), r.shape[0]),
                                "racc_cuda_scene_create_from_images"))


def create_environment(colors: np.ndarray) -> Environment:
    """racc::createEnvironment(context, colors, width, height) (RayAccelerator.h:111). colors: (H,W,4) float32."""
    c = np.ascontiguousarray(colors, dtype=np.float32)
    if c.ndim != 3 or c.shape[2] != 4:
        raise ValueError("colors must be (H, W, 4) float32 (racc::Color)")
    lib = _lib.load()
    return Environment(_lib.check_ptr(lib.racc_cuda_env_create(_ptr(c), c.shape[1], c.shape[0]), "racc_cuda_env_create"), c.shape[1], c.shape[0])


def _cuda_stream_handle(stream) -> ctypes.c_void_p:
    if stream is None:
        return ctypes.c_void_p(0)
    if hasattr(stream, "cuda_stream"):  # torch.cuda.Stream
        return ctypes.c_void_p(stream.cuda_stream)
    return ctypes.c_void_p(int(stream))


def trace_host_ptrs(scene: Scene, environment: Environment | None, streams, stream=None, counters_ptr: int | None = None) -> None:
    """HOST ray streams by raw pointer [(rays_ptr, results_ptr, count), ...] (e.g. pinned torch tensors).
    Asynchronous: complete after sync(stream)."""
    n = len(streams)
    arr = (StreamDesc * n)()
    for k, (rp, op, cnt) in enumerate(streams):
        arr[k] = StreamDesc(rp, op, cnt, _lib.STREAM_HOST)
    lib = _lib.load()
    h = _cuda_stream_handle(stream)
    env = environment._h if environment else None
    if counters_ptr is None:
        _lib.check(lib.racc_cuda_trace(scene._h, env, arr, n, h), "racc_cuda_trace")
    else:
        _lib.check(lib.racc_cuda_trace_counted(scene._h, env, arr, n, h, ctypes.c_void_p(counters_ptr), 0), "racc_cuda_trace_counted")


def trace_host(scene: Scene, environment: Environment | None, rays: np.ndarray, results: np.ndarray | None = None,
               stream=None, sync: bool = True) -> np.ndarray:
    """One ray stream in HOST memory through the engine: H2D, traversal, D2H (what gpuWorkerThread
    does per stream, RayAccelerator.cpp:369-410). rays: RAY_DTYPE array. Returns RESULT_DTYPE array."""
    rays = np.ascontiguousarray(rays)
    if rays.dtype != RAY_DTYPE:
        raise ValueError("rays must have RAY_DTYPE")
    if results is None:
        results = np.zeros(rays.shape[0], dtype=RESULT_DTYPE)
    elif results.dtype != RESULT_DTYPE or results.ndim != 1 or results.shape[0] < rays.shape[0] or not results.flags["C_CONTIGUOUS"]:
        raise ValueError("results must be a C-contiguous RESULT_DTYPE array with at least one record per ray "
                         "(the engine writes 16 bytes per ray into it)")
    desc = StreamDesc(rays.ctypes.data, results.ctypes.data, rays.shape[0], _lib.STREAM_HOST)
    lib = _lib.load()
    h = _cuda_stream_handle(stream)
    _lib.check(lib.racc_cuda_trace(scene._h, environment._h if environment else None, ctypes.byref(desc), 1, h), "racc_cuda_trace")
    if sync:
        _lib.check(lib.racc_cuda_sync(h), "racc_cuda_sync")
    return results


class PackedStreams:
    """A stream list already marshalled into the C-ABI's descriptor array (pack_streams): lets a caller with
    thousands of streams keep the Python loop out of a timed region."""

    def __init__(self, arr, n):
        self.arr, self.n = arr, n


def pack_streams(streams) -> PackedStreams:
    n = len(streams)
    arr = (StreamDesc * n)()
    for k, (rp, op, cnt) in enumerate(streams):
        arr[k] = StreamDesc(rp, op, cnt, _lib.STREAM_DEVICE)
    return PackedStreams(arr, n)


def trace_device(scene: Scene, environment: Environment | None, streams, stream=None, counters_ptr: int | None = None,
                 detail: bool = True) -> None:
    """Launch ONE traversal over a list of device-resident streams [(rays_ptr, results_ptr, count), ...]
    (or a PackedStreams). Asynchronous on `stream`. counters_ptr: device pointer to a zeroed Counters record
    (racc_cuda_counters: 8 x u64 -- rays, hits, inner nodes, pairs, stack pushes, leaf visits, 2 reserved); detail=False
    accumulates rays+hits only (free), detail=True also the visit counters (slower)."""
    packed = streams if isinstance(streams, PackedStreams) else pack_streams(streams)
    arr, n = packed.arr, packed.n
    lib = _lib.load()
    h = _cuda_stream_handle(stream)
    env = environment._h if environment else None
    if counters_ptr is None:
        _lib.check(lib.racc_cuda_trace(scene._h, env, arr, n, h), "racc_cuda_trace")
    else:
        _lib.check(lib.racc_cuda_trace_counted(scene._h, env, arr, n, h, ctypes.c_void_p(counters_ptr), 1 if detail else 0),
                   "racc_cuda_trace_counted")


def sync(stream=None) -> None:
    _lib.check(_lib.load().racc_cuda_sync(_cuda_stream_handle(stream)), "racc_cuda_sync")


def launch_count() -> int:
    return int(_lib.load().racc_cuda_launch_count())


def debug_warp_stats(reset: bool = True) -> list[int]:
    out = (ctypes.c_uint64 * 8)()
    _lib.check(_lib.load().racc_cuda_debug_warp_stats(out, 1 if reset else 0), "racc_cuda_debug_warp_stats")
    return [int(x) for x in out]


def set_tuning(**kw) -> None:
    keys = {"variant": 0, "block": 1, "ctas_per_sm": 2, "smem_nodes": 3, "fetch_threshold": 4, "leaf_bail": 5, "carveout": 6, "inner_bail": 7,
            "sort": 8, "sort_origin_bits": 9, "sort_dir_bits": 10, "sort_dir_major": 11, "build_device": 12, "smem_stack": 13, "host_zero_copy": 14,
            "whitted_arena": 15, "whitted_combine": 16, "host_taper": 17, "path_sync": 18, "path_stream": 19, "path_trace_ctas": 20}
    lib = _lib.load()
    unknown = [k for k in kw if k not in keys]
    if unknown:
        raise ValueError(f"unknown tuning key(s) {unknown}; known: {sorted(keys)}")
    for k, v in kw.items():
        lib.racc_cuda_set_tuning(keys[k], int(v))  # returns the previous value (which may itself be -1)


def generate_primary(camera: Camera, width: int, height: int, spp: int, jitter_seed: int, rays_ptr: int, stream=None) -> int:
    cam = _lib.CameraStruct()
    for k in range(3):
        cam.origin[k] = float(camera.origin[k]); cam.view[k] = float(camera.view[k])
        cam.right[k] = float(camera.right[k]); cam.up[k] = float(camera.up[k])
    _lib.check(_lib.load().racc_cuda_generate_primary(ctypes.byref(cam), width, height, spp, jitter_seed,
                                                       ctypes.c_void_p(rays_ptr), _cuda_stream_handle(stream)), "racc_cuda_generate_primary")
    return width * height * spp


def generate_bounce(scene: Scene, rays_ptr: int, results_ptr: int, count: int, seed: int, out_rays_ptr: int,
                    out_count_ptr: int, stream=None) -> None:
    _lib.check(_lib.load().racc_cuda_generate_bounce(scene._h, ctypes.c_void_p(rays_ptr), ctypes.c_void_p(results_ptr), count, seed,
                                                      ctypes.c_void_p(out_rays_ptr), ctypes.c_void_p(out_count_ptr),
                                                      _cuda_stream_handle(stream)), "racc_cuda_generate_bounce")


# ---- device-side wavefront path tracer (csrc/pathtrace.cu) ----------------------

# the four materials the reference assigns to battlefield.bin (Renderer/main.cpp:165-168): {r, g, b, eta}
BATTLEFIELD_MATERIALS = np.array([[0.8, 0.8, 0.8, 1.0 / 1.4], [0.1, 0.1, 0.1, 1.0 / 1.4], [0.6, 0.6, 0.6, 1.0 / 1.2], [0.3, 0.3, 0.3, 1.0 / 1.2]],
                                 dtype=np.float32)


class Shading:
    """Device copies of what the reference's example path tracer shades with (Renderer/SceneData.h:13-30)."""

    def __init__(self, handle):
        self._h = handle

    def destroy(self) -> None:
        if self._h:
            _lib.load().racc_cuda_shading_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def create_shading(normals: np.ndarray, triangle_normals: np.ndarray, triangle_materials: np.ndarray,
                   materials: np.ndarray = BATTLEFIELD_MATERIALS) -> Shading:
    """normals (V,4) float32, triangle_normals (T,4) float32, triangle_materials (T,) uint16, materials (M,4) float32 {r,g,b,eta}."""
    n = np.ascontiguousarray(normals, dtype=np.float32)
    tn = np.ascontiguousarray(triangle_normals, dtype=np.float32)
    tm = np.ascontiguousarray(triangle_materials, dtype=np.uint16).reshape(-1)
    m = np.ascontiguousarray(materials, dtype=np.float32)
    if n.ndim != 2 or n.shape[1] != 4 or tn.ndim != 2 or tn.shape[1] != 4 or m.ndim != 2 or m.shape[1] != 4:
        raise ValueError("normals, triangle_normals and materials must be (N, 4) float32")
    if tn.shape[0] != tm.shape[0]:
        raise ValueError("one geometric normal and one material index per triangle")
    d = _lib.ShadingDesc(n.ctypes.data, n.shape[0], tn.ctypes.data, tm.ctypes.data, tm.shape[0], m.ctypes.data, m.shape[0])
    return Shading(_lib.check_ptr(_lib.load().racc_cuda_shading_create(ctypes.byref(d)), "racc_cuda_shading_create"))


def _camera_struct(camera: Camera) -> "_lib.CameraStruct":
    cam = _lib.CameraStruct()
    for k in range(3):
        cam.origin[k] = float(camera.origin[k]); cam.view[k] = float(camera.view[k])
        cam.right[k] = float(camera.right[k]); cam.up[k] = float(camera.up[k])
    return cam


def path_trace(scene: Scene, environment: Environment | None, shading: Shading, camera: Camera, width: int, height: int, spp: int,
               max_depth: int, seed: int, framebuffer_ptr: int | None = None, sample_base: int = 0, batch_spp: int = 0, stream=None):
    """`spp` paths per pixel of the reference's path tracer (PathTracingRenderer.cpp) with the shading on the device.
    framebuffer_ptr: DEVICE pointer to width*height float4 radiance sums, added to in place; None = a fresh host array,
    returned as (H, W, 4) float32. Returns (framebuffer or None, rays traced per depth)."""
    return _render("racc_cuda_path_trace", scene, environment, shading, camera, width, height, spp, max_depth, seed, framebuffer_ptr,
                   sample_base, batch_spp, stream)


def whitted_trace(scene: Scene, environment: Environment | None, shading: Shading, camera: Camera, width: int, height: int, spp: int,
                  max_depth: int, seed: int, framebuffer_ptr: int | None = None, sample_base: int = 0, batch_spp: int = 0, stream=None):
    """`spp` samples per pixel of the reference's Whitted renderer (WhittedRenderer.cpp) with the shading on the device; arguments
    and return value as path_trace."""
    return _render("racc_cuda_whitted_trace", scene, environment, shading, camera, width, height, spp, max_depth, seed, framebuffer_ptr,
                   sample_base, batch_spp, stream)


def _render(entry: str, scene, environment, shading, camera, width, height, spp, max_depth, seed, framebuffer_ptr, sample_base, batch_spp, stream):
    waves = (ctypes.c_uint64 * (max_depth + 1))()
    d = _lib.PathDesc(width, height, sample_base, spp, max_depth, seed, batch_spp, 0)
    fb = None
    if framebuffer_ptr is None:
        fb = np.zeros((height, width, 4), dtype=np.float32)
        d.flags = _lib.FRAMEBUFFER_HOST
        framebuffer_ptr = fb.ctypes.data
    cam = _camera_struct(camera)
    _lib.check(getattr(_lib.load(), entry)(scene._h, environment._h if environment is not None else None, shading._h, ctypes.byref(cam),
                                           ctypes.byref(d), ctypes.c_void_p(framebuffer_ptr), waves, _cuda_stream_handle(stream)), entry)
    return fb, [int(x) for x in waves]
