"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

The path shards by timepoint: rank r owns one contiguous block of
timepoints, runs the whole pipeline on it, and takes part in exactly two exchanges:

  * `allreduce_maxima`  -- MAX over ranks of the two float64 flat-field maxima between pass 1 and
    pass 2 (both maxima are global over channel x time x tiles, preprocess.py:84,86);
  * `gather_summaries`  -- all-gather of the per-marker summaries (M, C, T_local, K) along time.

Images and ROI crops stay rank-local.  Works with the nccl backend on GPUs and with gloo on CPU
tensors (the world_size-2 tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_timepoints(num_times: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of timepoints owned by `rank`; blocks differ by at most one
    timepoint.  Contiguous (not strided) so that chip copy-forward timesteps, whose centres come
    from the nearest earlier search timestep (find.py:151), stay with their source when possible."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(num_times, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(num_times: int, world: int) -> List[int]:
    return [b - a for a, b in (shard_timepoints(num_times, r, world) for r in range(world))]


def allreduce_maxima(maxima: torch.Tensor, group=None) -> torch.Tensor:
    """In-place MAX all-reduce of the float64[2] flat-field maxima."""
    if maxima.dtype != torch.float64 or maxima.numel() != 2:
        raise ValueError("maxima must be a float64 tensor with 2 elements")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(maxima, op=dist.ReduceOp.MAX, group=group)
    return maxima


def gather_summaries(stats_local: torch.Tensor, num_times: int, group=None) -> torch.Tensor:
    """All-gather (M, C, T_local, K) summaries into (M, C, T, K) on every rank, T_local following
    `shard_timepoints`.  Uneven shards are padded to the largest one for the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats_local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(num_times, world)
    m, c, t_local, k = stats_local.shape
    if t_local != sizes[rank]:
        raise ValueError(f"rank {rank} holds {t_local} timepoints, expected {sizes[rank]}")
    t_max = max(sizes)
    send = stats_local
    if t_local != t_max:
        send = stats_local.new_zeros((m, c, t_max, k))
        send[:, :, :t_local] = stats_local
    send = send.contiguous()
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return torch.cat([r[:, :, :n] for r, n in zip(recv, sizes)], dim=2)


def broadcast_centres(centres: Optional[torch.Tensor], src: int = 0, group=None, device=None) -> torch.Tensor:
    """Broadcast the marker centres found on one rank (CPU finder) to all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return centres
    rank = dist.get_rank(group)
    shape = torch.tensor(list(centres.shape) if rank == src else [0, 0], dtype=torch.int64, device=device)
    dist.broadcast(shape, src=src, group=group)
    if rank != src:
        centres = torch.empty(tuple(int(v) for v in shape), dtype=torch.float64, device=device)
    dist.broadcast(centres, src=src, group=group)
    return centres


class SymmetricSummaries:
    """Gathered per-marker summaries (ranks, M, C, T, K) in peer-mapped (symmetric) memory.

    Every rank allocates the same buffer, the ranks exchange handles once (`rendezvous`), and the
    gather kernel of rank r stores its records directly into block r of EVERY rank's buffer over
    NVLink (`peer_blocks` are those block addresses, one per rank).  After `barrier()` the whole
    gathered tensor is valid on every rank: the all-gather of the summaries is fused into the
    kernel that computes them.  Raises if symmetric memory is unavailable; callers fall back to
    `gather_summaries` (NCCL all-gather)."""

    def __init__(self, shape_local, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.shape_local = tuple(int(s) for s in shape_local)
        self.gathered = symm_mem.empty((self.world,) + self.shape_local, dtype=torch.float64, device=device)
        self.handle = symm_mem.rendezvous(self.gathered, group)
        block_bytes = 8
        for s in self.shape_local:
            block_bytes *= s
        self.peer_blocks = [int(ptr) + self.rank * block_bytes for ptr in self.handle.buffer_ptrs]

    def barrier(self) -> None:
        """Cross-rank barrier on the current stream: every rank's stores have landed everywhere."""
        self.handle.barrier()
