"""ctypes loader for libracc_b200.so (the C-ABI of include/racc_b200.h).

There is no Python or CPU implementation behind this module: if the shared library is missing
or a CUDA device is absent, the calls fail loudly.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libracc_b200.so")

STREAM_DEVICE = 0
STREAM_HOST = 1


class StreamDesc(ctypes.Structure):
    _fields_ = [("rays", ctypes.c_void_p), ("results", ctypes.c_void_p), ("count", ctypes.c_uint32), ("flags", ctypes.c_uint32)]


class SceneInfo(ctypes.Structure):
    _fields_ = [
        ("node_count", ctypes.c_uint32), ("pair_count", ctypes.c_uint32), ("real_pair_count", ctypes.c_uint32),
        ("remap_count", ctypes.c_uint32), ("depth", ctypes.c_uint32), ("triangle_count", ctypes.c_uint32),
        ("bounds_min", ctypes.c_float * 3), ("bounds_max", ctypes.c_float * 3),
    ]


class Counters(ctypes.Structure):
    _fields_ = [("rays", ctypes.c_uint64), ("hits", ctypes.c_uint64), ("inner_nodes", ctypes.c_uint64), ("pairs_tested", ctypes.c_uint64),
                ("stack_pushes", ctypes.c_uint64), ("leaf_visits", ctypes.c_uint64), ("reserved", ctypes.c_uint64 * 2)]


class CameraStruct(ctypes.Structure):
    _fields_ = [("origin", ctypes.c_float * 3), ("view", ctypes.c_float * 3), ("right", ctypes.c_float * 3), ("up", ctypes.c_float * 3)]


class ShadingDesc(ctypes.Structure):
    _fields_ = [("normals4", ctypes.c_void_p), ("vertex_count", ctypes.c_uint32), ("triangle_normals4", ctypes.c_void_p),
                ("triangle_materials", ctypes.c_void_p), ("triangle_count", ctypes.c_uint32), ("materials_ke4", ctypes.c_void_p),
                ("material_count", ctypes.c_uint32)]


class PathDesc(ctypes.Structure):
    _fields_ = [("width", ctypes.c_uint32), ("height", ctypes.c_uint32), ("sample_base", ctypes.c_uint32), ("spp", ctypes.c_uint32),
                ("max_depth", ctypes.c_uint32), ("seed", ctypes.c_uint32), ("batch_spp", ctypes.c_uint32), ("flags", ctypes.c_uint32)]


FRAMEBUFFER_HOST = 1

# every symbol include/racc_b200.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_U32 = ctypes.c_uint32
SYMBOLS = {
    "racc_cuda_init": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "racc_cuda_device_count": (ctypes.c_int, []),
    "racc_cuda_current_devices": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "racc_cuda_thread_release": (None, []),
    "racc_cuda_frame_reduce": (ctypes.c_int, [ctypes.POINTER(Counters), _P]),
    "racc_cuda_comm_unique_id": (ctypes.c_int, [_P]),
    "racc_cuda_comm_init_rank": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int]),
    "racc_cuda_comm_destroy": (None, []),
    "racc_cuda_comm_ranks": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "racc_cuda_gather_results": (ctypes.c_int, [_P, _U32, _P, _P]),
    "racc_cuda_abi_version": (ctypes.c_int, []),
    "racc_cuda_last_error": (ctypes.c_char_p, []),
    "racc_cuda_scene_create": (_P, [_P, _U32, _P, _U32]),
    "racc_cuda_scene_create_from_images": (_P, [_P, _U32, _P, _U32, _P, _U32]),
    "racc_cuda_build_images": (_P, [_P, _U32, _P, _U32]),
    "racc_cuda_host_images_get_info": (ctypes.c_int, [_P, ctypes.POINTER(SceneInfo)]),
    "racc_cuda_host_images_copy": (ctypes.c_int, [_P, _P, _P, _P]),
    "racc_cuda_host_images_destroy": (None, [_P]),
    "racc_cuda_scene_destroy": (None, [_P]),
    "racc_cuda_scene_get_info": (ctypes.c_int, [_P, ctypes.POINTER(SceneInfo)]),
    "racc_cuda_scene_download": (ctypes.c_int, [_P, _P, _P, _P]),
    "racc_cuda_env_create": (_P, [_P, _U32, _U32]),
    "racc_cuda_env_destroy": (None, [_P]),
    "racc_cuda_trace": (ctypes.c_int, [_P, _P, ctypes.POINTER(StreamDesc), _U32, _P]),
    "racc_cuda_trace_counted": (ctypes.c_int, [_P, _P, ctypes.POINTER(StreamDesc), _U32, _P, _P, ctypes.c_int]),
    "racc_cuda_host_alloc": (_P, [ctypes.c_size_t]),
    "racc_cuda_host_free": (None, [_P]),
    "racc_cuda_stream_create": (_P, []),
    "racc_cuda_stream_destroy": (None, [_P]),
    "racc_cuda_sync": (ctypes.c_int, [_P]),
    "racc_cuda_launch_count": (ctypes.c_uint64, []),
    "racc_cuda_debug_rcp_table": (ctypes.c_int, [ctypes.POINTER(ctypes.c_float)]),
    "racc_cuda_debug_warp_stats": (ctypes.c_int, [ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]),
    "racc_cuda_set_variant": (ctypes.c_int, [ctypes.c_int]),
    "racc_cuda_set_tuning": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "racc_cuda_generate_primary": (ctypes.c_int, [ctypes.POINTER(CameraStruct), _U32, _U32, _U32, _U32, _P, _P]),
    "racc_cuda_generate_bounce": (ctypes.c_int, [_P, _P, _P, _U32, _U32, _P, _P, _P]),
    "racc_cuda_shading_create": (_P, [ctypes.POINTER(ShadingDesc)]),
    "racc_cuda_shading_destroy": (None, [_P]),
    "racc_cuda_path_trace": (ctypes.c_int, [_P, _P, _P, ctypes.POINTER(CameraStruct), ctypes.POINTER(PathDesc), _P,
                                             ctypes.POINTER(ctypes.c_uint64), _P]),
    "racc_cuda_whitted_trace": (ctypes.c_int, [_P, _P, _P, ctypes.POINTER(CameraStruct), ctypes.POINTER(PathDesc), _P,
                                                ctypes.POINTER(ctypes.c_uint64), _P]),
}

_lib = None


class EngineError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Loads the engine. Raises if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python -m rayaccel_b200.build` "
            "(or __graft_entry__.build()). rayaccel_b200 has no CPU or pure-Python fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().racc_cuda_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise EngineError(f"{what}: {last_error()}")


def check_ptr(ptr, what: str):
    if not ptr:
        raise EngineError(f"{what}: {last_error()}")
    return ptr
