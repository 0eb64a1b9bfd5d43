"""How long does the HOST take to enqueue one racc_cuda_path_trace frame (no wave counts asked: the call does not wait), against
the GPU time of that frame? Development tool.  usage: path_enqueue_time.py [width height spp]"""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402
from rayaccel_b200 import _lib, api, scene_io  # noqa: E402

w, h, spp = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (1920, 1080, 1)))
torch.cuda.set_device(0)
rb.init(0)
sf = rb.load_scene()
scene = rb.create_scene(sf.vertices, sf.indices)
env = rb.create_environment(sf.environment)
shading = rb.create_shading(sf.normals, sf.triangle_normals, sf.materials)
cam = api._camera_struct(scene_io.Camera.look_at(sf.cam_origin, sf.cam_target, sf.cam_up, sf.cam_fov, w, h))
fb = torch.zeros(w * h * 4, dtype=torch.float32, device="cuda")
d = _lib.PathDesc(w, h, 0, spp, 3, 1, 0, 0)
fn = _lib.load().racc_cuda_path_trace
for waves in (False, True):
    host, gpu = [], []
    wv = (ctypes.c_uint64 * 4)()
    for rep in range(12):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        t0 = time.perf_counter()
        rc = fn(scene._h, env._h, shading._h, ctypes.byref(cam), ctypes.byref(d), ctypes.c_void_p(fb.data_ptr()), wv if waves else None, None)
        t1 = time.perf_counter()
        b.record()
        torch.cuda.synchronize()
        assert rc == 0
        if rep >= 2:
            host.append((t1 - t0) * 1e3)
            gpu.append(a.elapsed_time(b))
    print(f"{w}x{h}x{spp} wave counts {'asked' if waves else 'not asked'}: host call {min(host):.3f} ms (median {sorted(host)[len(host)//2]:.3f}), events {min(gpu):.3f} ms")
