"""TEST HELPER, launched under torchrun by tests/test_gpu_multi.py: one process per GPU, the engine's own communicator
(racc_cuda_comm_unique_id / _init_rank), a frame whose rays are sharded by index, the per-frame hit reduction and the
all-gather of the Result slices -- both NCCL calls made by the engine library. Rank 0 checks the gathered, index-parallel
hit buffer against the CPU oracle on the whole frame and prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
import rayaccel_b200 as rb  # noqa: E402
from conftest import random_rays  # noqa: E402
from rayaccel_b200 import sharding  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group(backend="gloo")  # only to hand the NCCL id round; the collectives below are the engine's
rb.init(local)
box = [rb.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
rb.comm_init_rank(box[0], rank, world)
assert rb.comm_ranks() == (rank, world)

sf = rb.load_scene()
scene = rb.create_scene(sf.vertices, sf.indices)
env = rb.create_environment(sf.environment)
total = 1_000_003  # not a multiple of the world size: the last slice is padded
rays = random_rays(total, sf.vertices[:, :3].min(0), sf.vertices[:, :3].max(0), seed=77)
per_rank = -(-total // world)
begin, end = rank * per_rank, min(total, (rank + 1) * per_rank)
mine = np.zeros(per_rank, dtype=oracle.RAY_DTYPE)
mine[: end - begin] = rays[begin:end]
mine["maxT"][end - begin:] = -1.0  # padding rays: empty interval, always a miss
d_rays = torch.from_numpy(mine.view(np.float32).reshape(-1)).cuda()
d_all = torch.zeros(world * per_rank * 4, dtype=torch.float32, device="cuda")
slice_ptr = d_all.data_ptr() + rank * per_rank * 16  # traced straight into this rank's place of the full buffer
rb.frame_reduce()
rb.trace_device(scene, env, [(d_rays.data_ptr(), slice_ptr, per_rank)])
rb.gather_results(slice_ptr, per_rank, d_all.data_ptr())
frame = rb.frame_reduce()
torch.cuda.synchronize()
got = d_all.cpu().numpy().view(np.uint32).reshape(-1, 4)[:total]
nodes, pairs, remap = scene.download()
want = oracle.traverse(oracle.SceneImages(nodes, pairs, remap, sf.environment), rays)
ok = bool(np.array_equal(got, want.view(np.uint32).reshape(-1, 4)))
hits = int((want["triangle"] != oracle.INVALID).sum())
flags = [None] * world
dist.all_gather_object(flags, ok)
if rank == 0:
    print(json.dumps({"ranks": world, "every_rank_holds_the_full_hit_buffer_bit_exact": all(flags), "frame_rays": frame["rays"],
                      "frame_rays_expected": world * per_rank, "frame_hits": frame["hits"], "frame_hits_expected": hits}))
rb.comm_destroy()
dist.destroy_process_group()
