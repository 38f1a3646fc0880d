"""Multi-GPU plumbing: one process per GPU, query cells sharded by rows, reference replicated.

The hot path has exactly one exchange step: the per-cluster MoE sufficient
statistics `gram [K,B1,B1]` / `cross [K,B1,d]` (float64, 0.05-10 MB) are summed over ranks between
the accumulate and solve kernels.  Everything else is independent per query cell, so there is no
other data-path collective; the batch-level vocabulary (a few strings) is agreed on with an
object all-gather on the host.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size() -> int:
    return dist.get_world_size() if active() else 1


def rank() -> int:
    return dist.get_rank() if active() else 0


def allreduce_stats(gram: torch.Tensor, cross: torch.Tensor) -> None:
    """In-place sum over ranks (NCCL on CUDA tensors, gloo on CPU tensors in the tests)."""
    if not active():
        return
    K, B1, _ = gram.shape
    flat = torch.cat([gram.reshape(K, -1), cross.reshape(K, -1)], dim=1).contiguous()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    gram.copy_(flat[:, : B1 * B1].reshape(gram.shape))
    cross.copy_(flat[:, B1 * B1:].reshape(cross.shape))


def broadcast(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    """`t` as rank `src` holds it, on every rank (in place; no-op outside a process group)."""
    if active():
        dist.broadcast(t, src=src)
    return t


def union_levels(local_levels: list[list[str]]) -> list[list[str]]:
    """Sorted union, per batch variable, of the level names seen on every rank."""
    if not active():
        return local_levels
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, local_levels)
    out = []
    for v in range(len(local_levels)):
        s = set()
        for g in gathered:
            s.update(g[v])
        out.append(sorted(s))
    return out
