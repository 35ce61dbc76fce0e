"""Object -> GPU partition of the multi-object NeRF backend, and the cross-rank timing reduction of the bench.

Objects are independent (no parameter, gradient or activation is shared between two nerf::NeRF_Model instances,
MON/Core/src/nerf_manager.cu:75-89), so the partition is the reference's round-robin rule `gpu = id mod #GPUs`
(MON/Core/src/nerf.cu:27-33) and there is NO data-path collective: ranks only meet at the barrier that brackets
the timed region and in the max-reduction of the per-rank device times.
"""
from __future__ import annotations


def assign_objects(n_objects: int, world_size: int) -> list[list[int]]:
    """objects[k] -> rank k % world_size; returns, per rank, the ascending list of object ids it trains."""
    if n_objects < 0 or world_size < 1:
        raise ValueError("n_objects >= 0 and world_size >= 1 required")
    return [list(range(r, n_objects, world_size)) for r in range(world_size)]


def owner(object_id: int, world_size: int) -> int:
    return object_id % world_size


def reduce_max(value: float, device=None) -> float:
    """max over ranks of a per-rank scalar (device time of the timed region); identity when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_all(value: float, device=None) -> list:
    """the per-rank scalars of all ranks, in rank order (diagnostics: which rank was the slowest); [value] when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(value)]
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


def reduce_sum(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
