"""Multi-GPU plumbing of the collision query path (one process per GPU, torch.distributed).

The path shards by space: each rank owns one Morton-rank slab of the query leaves (the C ABI's
``c2d_gpu_set_shard(rank, world)`` uses :func:`slab_bounds`) and emits only the pairs whose query leaf it owns;
rank 0 gathers the variable-size lists of 84-byte ``EntityCollision`` records.  Works with NCCL (GPU tensors)
and gloo (CPU tensors, used by the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def slab_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of slab `rank`: the same integer arithmetic as launch_traverse() in csrc/c2d_gpu.cu."""
    return (n * rank) // world, (n * (rank + 1)) // world


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce_sum(value: float) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_max(value: float) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_records(records: np.ndarray, dst: int = 0) -> np.ndarray | None:
    """Variable-size gather of a structured record array to rank `dst` (all-gather of the counts, then padded
    byte tensors).  Returns the concatenation on `dst`, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return records
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _device()
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = len(records)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    item = records.dtype.itemsize
    width = int(counts.max().item()) * item
    buf = torch.zeros(max(width, 1), dtype=torch.uint8, device=dev)
    if len(records):
        raw = torch.from_numpy(np.frombuffer(records.tobytes(), dtype=np.uint8).copy())
        buf[: raw.numel()] = raw.to(dev)
    parts = [torch.zeros_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst)
    if rank != dst:
        return None
    out = [np.frombuffer(parts[r][: int(counts[r].item()) * item].cpu().numpy().tobytes(), dtype=records.dtype)
           for r in range(world)]
    return np.concatenate(out) if out else records[:0]
