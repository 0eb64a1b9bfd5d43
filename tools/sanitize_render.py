"""Small device-side path-tracing frame for compute-sanitizer (memcheck / racecheck): two lanes, three batches (the last
one ragged), device and host framebuffers; the two must agree."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402
from rayaccel_b200 import scene_io  # noqa: E402

torch.cuda.set_device(0)
rb.init(0)
sf = rb.load_scene()
scene = rb.create_scene(sf.vertices, sf.indices)
env = rb.create_environment(sf.environment)
shading = rb.create_shading(sf.normals, sf.triangle_normals, sf.materials)
w, h = 256, 128
cam = scene_io.Camera.look_at(sf.cam_origin, sf.cam_target, sf.cam_up, sf.cam_fov, w, h)
host, waves = rb.path_trace(scene, env, shading, cam, w, h, 9, 3, seed=5, batch_spp=4)  # 131072 paths per batch: two lanes
fb = torch.zeros(w * h * 4, dtype=torch.float32, device="cuda")
_, waves2 = rb.path_trace(scene, env, shading, cam, w, h, 9, 3, seed=5, framebuffer_ptr=fb.data_ptr(), batch_spp=3)  # one lane
torch.cuda.synchronize()
assert waves == waves2 and fb.cpu().numpy().tobytes() == host.tobytes(), "host and device framebuffers differ"
print("sanitize render ok:", sum(waves), "rays", waves)
# the same frames as one persistent kernel per batch (tuning key 19, pathstream.cu): queue, tickets, epoch flags
rb.set_tuning(path_stream=1)
streamed, waves3 = rb.path_trace(scene, env, shading, cam, w, h, 9, 3, seed=5, batch_spp=4)
streamed2, waves4 = rb.path_trace(scene, env, shading, cam, w, h, 9, 3, seed=5, batch_spp=3)
rb.set_tuning(path_stream=0)
assert waves3 == waves and waves4 == waves and streamed.tobytes() == host.tobytes() and streamed2.tobytes() == host.tobytes(), "streamed form differs"
print("sanitize streamed render ok:", sum(waves3), "rays")
