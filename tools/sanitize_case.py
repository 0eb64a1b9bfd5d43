"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck): device scene build, packed
kernel with the shared-memory stack tops and ray re-binning, reference-format kernel with TMA-staged nodes.
Checks the results against each other (not against the oracle: the sanitizer run is about memory safety)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rayaccel_b200 as rb  # noqa: E402

torch.cuda.set_device(0)
rb.init(0)
n_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
v, i = rb.synthetic_triangles(n_tris, seed=3, extent=60.0, edge=3.0)
scenes = []
for mode in (0, 2):
    rb.set_tuning(build_device=mode)
    scenes.append(rb.create_scene(v, i))
rb.set_tuning(build_device=3)
a, b = scenes[0].download(), scenes[1].download()
assert all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip(a, b)), "device build differs from host build"
rng = np.random.default_rng(5)
n = 60000
rays = np.zeros((n, 8), dtype=np.float32)
rays[:, 0:3] = rng.uniform(-5, 65, size=(n, 3))
d = rng.normal(size=(n, 3))
rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
rays[:, 7] = 1e6
d_rays = torch.from_numpy(rays.reshape(-1)).cuda()
outs = []
for tuning in (dict(variant=3, sort=0, smem_stack=0), dict(variant=3, sort=1, smem_stack=16), dict(variant=3, sort=1, smem_stack=8, sort_dir_bits=3),
               dict(variant=2, smem_nodes=-1), dict(variant=0, smem_nodes=64), dict(variant=1)):
    rb.set_tuning(**{**dict(variant=3, block=256, ctas_per_sm=5, smem_nodes=0, sort=0, smem_stack=0, sort_dir_bits=0), **tuning})
    o = torch.zeros(n * 4, dtype=torch.float32, device="cuda")
    cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
    rb.trace_device(scenes[1], None, [(d_rays.data_ptr(), o.data_ptr(), n // 2), (d_rays[(n // 2) * 8:].data_ptr(), o[(n // 2) * 4:].data_ptr(), n - n // 2)],
                    counters_ptr=cnt.data_ptr(), detail=True)
    torch.cuda.synchronize()
    outs.append(o.view(torch.int32).cpu())
assert all(torch.equal(outs[0], o) for o in outs[1:]), "kernel variants disagree"
# round 2: the quantised-node kernel (both stack flavours, re-binned), HOST streams through the staging pipeline with the tapered
# tail, the frame record, and thread teardown
for tuning in (dict(variant=4), dict(variant=4, sort=1, smem_stack=16)):
    rb.set_tuning(**{**dict(variant=3, block=256, ctas_per_sm=5, smem_nodes=0, sort=0, smem_stack=0, sort_dir_bits=0), **tuning})
    o = torch.zeros(n * 4, dtype=torch.float32, device="cuda")
    rb.trace_device(scenes[1], None, [(d_rays.data_ptr(), o.data_ptr(), n)])
    torch.cuda.synchronize()
    ids = o.view(torch.int32).cpu().view(-1, 4)[:, 0]
    assert int((ids != outs[0].view(-1, 4)[:, 0]).sum()) <= 2, "quantised nodes: ids differ beyond ties"
rb.set_tuning(variant=3, sort=2, smem_stack=-1)
host_rays = torch.from_numpy(rays.reshape(-1)).pin_memory()
host_out = torch.zeros(n * 4, dtype=torch.float32).pin_memory()
rb.frame_reduce()
rb.trace_host_ptrs(scenes[1], None, [(host_rays.data_ptr(), host_out.data_ptr(), 7), (host_rays.data_ptr() + 7 * 32, host_out.data_ptr() + 7 * 16, n - 7)])
rb.sync()
assert torch.equal(host_out.view(torch.int32), outs[0]), "HOST stream results differ"
assert rb.frame_reduce()["rays"] == n
rb.thread_release()
print("sanitize case ok:", n_tris, "triangles,", n, "rays,", int((outs[0].view(-1, 4)[:, 0] != -1).sum()), "hits")
