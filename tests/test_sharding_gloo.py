"""The N>1 path on CPU: world_size-2 `gloo` processes shard a ray set by index, each "traces" its
slice (the oracle stands in for the device here -- this test is about the host-side sharding,
gather and reduction logic in rayaccel_b200/sharding.py), then the gathered results and reduced
frame counters must equal a single-process run over the whole set."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from rayaccel_b200 import sharding  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 8, 1000003, 8294400):
        for world in (1, 2, 3, 4, 8):
            b = [sharding.shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = sharding.shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total
    assert sharding.deal_streams(5, 1, 2) == [1, 3]
    assert sorted(sum((sharding.deal_streams(11, r, 4) for r in range(4)), [])) == list(range(11))
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import oracle
    import rayaccel_b200 as rb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.load(os.path.join(GOLDEN, "battlefield_rays.npz"))
        rays = np.ascontiguousarray(g["rays"][:total]).view(oracle.RAY_DTYPE).reshape(-1)
        sf = rb.load_scene()
        h = rb.HostImages(sf.vertices, sf.indices)  # scene replicated: every rank builds the same images
        images = oracle.SceneImages(h.nodes, h.pairs, h.remap, sf.environment)
        lo, hi = sharding.shard_bounds(total, rank, world)
        res, cnt = oracle.traverse(images, rays[lo:hi], counters=True, threads=2)
        local = torch.from_numpy(res.view(np.float32).reshape(-1).copy())
        full = sharding.gather_results(local, total)
        counters = torch.tensor([hi - lo, int(cnt["hit"].sum()), int(cnt["inner"].astype(np.int64).sum()), int(cnt["pairs"].astype(np.int64).sum())],
                                dtype=torch.int64)
        sharding.reduce_frame_counters(counters)
        # the load-balanced policy: runs of 500 rays dealt round-robin, traced run by run, gathered back in index order
        runs = sharding.interleaved_blocks(total, rank, world, 500)
        parts = [oracle.traverse(images, rays[b:e], threads=2).view(np.float32).reshape(-1) for b, e in runs]
        local_i = torch.from_numpy(np.concatenate(parts).copy()) if parts else torch.zeros(0)
        full_i = sharding.gather_results_interleaved(local_i, total, 500)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), full=full.numpy().view(np.uint32).reshape(-1, 4), counters=counters.numpy(),
                 full_interleaved=full_i.numpy().view(np.uint32).reshape(-1, 4))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [6144, 4097])  # even and ragged split
def test_two_rank_sharding_matches_single_process(tmp_path, total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    g = np.load(os.path.join(GOLDEN, "battlefield_rays.npz"))
    want = g["results"][:total]
    for rank in range(world):
        out = np.load(tmp_path / f"rank{rank}.npz")
        assert np.array_equal(out["full"], want), f"rank {rank}: gathered results differ from the single-process golden run"
        assert np.array_equal(out["full_interleaved"], want), f"rank {rank}: interleaved gather differs from the golden run"
        assert out["counters"][0] == total
        assert out["counters"][1] == int((want[:, 0] != 0xFFFFFFFF).sum())
        assert out["counters"][2] == int(g["inner"][:total].astype(np.int64).sum())
        assert out["counters"][3] == int(g["pairs"][:total].astype(np.int64).sum())
