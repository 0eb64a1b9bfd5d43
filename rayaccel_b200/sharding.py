"""Multi-GPU sharding of ray streams: one process per GPU, scene replicated, rays split by index.

The reference treats any stream as testable by any back-end (RayAccelerator.cpp:287-296,356-363) and
rays carry no inter-ray state (Kernels.h:142), so the path shards with NO data-path collective: each
rank traces its own contiguous slice (or its own whole streams). The only exchange is per frame:
an all-reduce of the frame counters {rays, hits, inner-node visits, pair tests} (what racc::Stats
reports and the roofline accounting uses) and -- only when one rank needs every hit -- an all-gather
of the 16-byte Result slices. Both go through torch.distributed (NCCL over NVLink/NVSwitch on the
GPUs, gloo in the CPU tests); tensors stay wherever the backend wants them.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [begin, end) of `total` rays owned by `rank`: sizes differ by at most one ray."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return total * rank // world, total * (rank + 1) // world


def shard_sizes(total: int, world: int) -> list[int]:
    return [shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0] for r in range(world)]


def deal_streams(n_streams: int, rank: int, world: int) -> list[int]:
    """Whole streams dealt round-robin (the analogue of gpuSubmissionThreads each popping streams)."""
    return list(range(rank, n_streams, world))


def reduce_frame_counters(counters: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the per-rank frame counters (int64 x4: rays, hits, inner, pairs) over all ranks, in place."""
    if counters.dtype != torch.int64 or counters.numel() != 4:
        raise ValueError("frame counters are 4 x int64")
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def gather_results(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather the ranks' Result slices (float32 words, 4 per ray) into the full, index-parallel
    result array of `total` rays on every rank. Slices follow shard_bounds, so they may be ragged by
    one ray; the shorter ones are padded for the collective and trimmed afterwards."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sizes = shard_sizes(total, world)
    if local.numel() != sizes[rank] * 4:
        raise ValueError(f"rank {rank} holds {local.numel() // 4} results, its shard has {sizes[rank]}")
    if world == 1:
        return local.clone()
    longest = max(sizes)
    padded = torch.zeros(longest * 4, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    gathered = torch.empty(world * longest * 4, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    if all(s == longest for s in sizes):
        return gathered
    return torch.cat([gathered[r * longest * 4: r * longest * 4 + sizes[r] * 4] for r in range(world)])


def interleaved_blocks(total: int, rank: int, world: int, block: int) -> list[tuple[int, int]]:
    """Load-balanced alternative to shard_bounds: the index range is cut into runs of `block` rays and run b
    belongs to rank b % world (a frame's sky rows are cheap and its ground rows expensive, so contiguous
    slices of a pixel-ordered stream are unbalanced). Returns this rank's runs as [begin, end) pairs."""
    if block <= 0 or not (0 <= rank < world):
        raise ValueError("bad block size or rank")
    return [(b, min(b + block, total)) for b in range(rank * block, total, world * block)]


def gather_results_interleaved(local: torch.Tensor, total: int, block: int, group=None) -> torch.Tensor:
    """All-gather for interleaved_blocks: `local` holds this rank's runs back to back (4 float32 words per
    ray); returns the full index-parallel result array on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    counts = [sum(e - b for b, e in interleaved_blocks(total, r, world, block)) for r in range(world)]
    if local.numel() != counts[rank] * 4:
        raise ValueError(f"rank {rank} holds {local.numel() // 4} results, its runs hold {counts[rank]}")
    if world == 1:
        return local.clone()
    longest = max(counts)
    padded = torch.zeros(longest * 4, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    gathered = torch.empty(world * longest * 4, dtype=local.dtype, device=local.device).view(world, longest, 4)
    dist.all_gather_into_tensor(gathered.view(-1), padded, group=group)
    out = torch.empty(total, 4, dtype=local.dtype, device=local.device)
    n_blocks = (total + block - 1) // block
    for r in range(world):
        # rank r's k-th run is global block r + k*world
        mine = torch.arange(r, n_blocks, world, device=local.device)
        if mine.numel() == 0:
            continue
        idx = (mine[:, None] * block + torch.arange(block, device=local.device)[None, :]).reshape(-1)
        idx = idx[idx < total]
        out[idx] = gathered[r, : idx.numel()]
    return out.view(-1)


def sample_range(spp: int, rank: int, world: int) -> tuple[int, int]:
    """Device-side renderer across GPUs: the frame's samples are split, every rank renders all pixels for
    samples [first, first + count) (racc_cuda_path_desc.sample_base / spp) and the framebuffers are summed."""
    begin, end = shard_bounds(spp, rank, world)
    return begin, end - begin


def reduce_framebuffer(framebuffer: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the ranks' radiance sums (float32, width*height*4) over all ranks, in place: the one collective of the
    device-side renderer (the reference accumulates one sample per frame into one host framebuffer,
    PathTracingRenderer.cpp:540-543). The sum is associative up to float rounding, so an N-GPU image equals the
    1-GPU image to a few ulp, not bit for bit."""
    if framebuffer.dtype != torch.float32:
        raise ValueError("the framebuffer holds float32 radiance sums")
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(framebuffer, op=dist.ReduceOp.SUM, group=group)
    return framebuffer
