"""rayaccel_b200 -- B200-native wavefront ray-intersection engine behind the RayAccelerator API.

The package holds only the hot path of rasmusbarr/rayaccel: scene images (BVH2 + triangle pairs),
the sm_100a traversal kernels and the C-ABI (csrc/, include/racc_b200.h), plus this thin Python
mirror of the reference's interface. There is no CPU fallback.
"""
from .api import (BATTLEFIELD_MATERIALS, INVALID_TRIANGLE, RAY_DTYPE, RESULT_DTYPE, Environment, HostImages, Scene, Shading,  # noqa: F401
                  comm_destroy, comm_init_rank, comm_ranks, comm_unique_id, gather_results, create_environment, create_scene, current_devices, frame_reduce, thread_release, create_scene_from_images, create_shading, debug_warp_stats, device_count, generate_bounce, generate_primary, init,
                  launch_count, pack_streams, path_trace, set_tuning, sync, trace_device, trace_host, trace_host_ptrs, whitted_trace)
from ._lib import EngineError  # noqa: F401
from .scene_io import Camera, SceneFile, load_scene, synthetic_triangles  # noqa: F401
