"""Application-level multi-GPU helpers of round 1 (one process per GPU, torch.distributed), kept for the tests that
exercise the ghost-copy owner rule and the record gather.  The decomposition itself now lives INSIDE the library
(include/c2d_gpu.h "multi-GPU": c2d_gpu_comm_init, Morton slabs, move exchange, result gather) - bench.py uses that.

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


class _DeviceBytes:
    """Wraps a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def gather_device_records(ptr: int, count: int, itemsize: int, dst: int = 0, pinned: dict | None = None):
    """GPU-to-GPU gather (NCCL) of each rank's `count` records of `itemsize` bytes living at device address `ptr`.
    Rank `dst` gets (host uint8 array of all records, total count) — downloaded once into a reusable pinned buffer;
    the other ranks get (None, total count)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = count
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    host_counts = counts.cpu().tolist()
    width = max(host_counts) * itemsize
    send = torch.zeros(max(width, 1), dtype=torch.uint8, device=dev)
    if count:
        send[: count * itemsize] = torch.as_tensor(_DeviceBytes(ptr, count * itemsize), device=dev)
    parts = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, parts, dst=dst)
    total = int(sum(host_counts))
    if rank != dst:
        return None, total
    packed = torch.cat([parts[r][: host_counts[r] * itemsize] for r in range(world)])
    if pinned is not None:
        buf = pinned.get("buf")
        if buf is None or buf.numel() < packed.numel():
            buf = torch.empty(max(packed.numel(), 1) * 2, dtype=torch.uint8, pin_memory=True)
            pinned["buf"] = buf
        out = buf[: packed.numel()]
        out.copy_(packed, non_blocking=False)
        return out.numpy(), total
    return packed.cpu().numpy(), total


# ---------------------------------------------------------------------------------------------------------------
# Spatial slabs with halo ("ghost") copies — the weak-scaling decomposition of a large world
# ---------------------------------------------------------------------------------------------------------------
def slab_edges(x: np.ndarray, world: int) -> np.ndarray:
    """world + 1 x-coordinates cutting the shapes into `world` slabs of (almost) equal population (quantiles, so a
    clustered world is balanced too); edges[0] = -inf, edges[-1] = +inf."""
    order = np.sort(np.asarray(x, np.float64))
    cuts = [order[(len(order) * r) // world] for r in range(1, world)]
    return np.array([-np.inf, *cuts, np.inf])


def slab_selection(x: np.ndarray, rank: int, world: int, margin: float):
    """(owned mask, ghost mask) of slab `rank`: owned = centre in [edge_r, edge_r+1); ghosts = centres of the other
    slabs within `margin` of this slab.  `margin` must exceed the largest interaction distance (fat box extents)
    plus the drift accumulated while the decomposition is kept."""
    e = slab_edges(x, world)
    lo, hi = e[rank], e[rank + 1]
    owned = (x >= lo) & (x < hi)
    near = (x >= lo - margin) & (x < hi + margin)
    return owned, near & ~owned


def slab_exports(x: np.ndarray, rank: int, world: int, margin: float) -> np.ndarray:
    """Mask of the shapes OWNED by `rank` that some neighbour holds as a halo copy (within `margin` of a slab edge)."""
    e = slab_edges(x, world)
    lo, hi = e[rank], e[rank + 1]
    owned = (x >= lo) & (x < hi)
    return owned & ((x < lo + margin) | (x >= hi - margin))


def exchange_boundary_moves(ids: np.ndarray, dxy: np.ndarray):
    """The per-frame exchange step of the slab decomposition: every rank contributes the moves (entity id, dx, dy) of
    its boundary shapes; everybody receives all of them (one all-gather of the counts, one padded all-gather of the
    12-byte records — NCCL over NVLink on GPUs, gloo in the CPU tests) and applies the ones that match its halo copies.
    Returns (ids, dxy) of all ranks concatenated."""
    ids = np.ascontiguousarray(ids, np.uint32)
    dxy = np.ascontiguousarray(dxy, np.float32).reshape(-1, 2)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return ids, dxy
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = _device()
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = len(ids)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    host_counts = counts.cpu().tolist()
    width = max(max(host_counts), 1)
    send = torch.zeros((width, 3), dtype=torch.float32, device=dev)
    if len(ids):
        rec = np.empty((len(ids), 3), np.float32)
        rec[:, 0] = ids.view(np.float32)  # bit pattern of the id travels in a float lane
        rec[:, 1:] = dxy
        send[: len(ids)] = torch.from_numpy(rec).to(dev)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    parts = [recv[r][: host_counts[r]].cpu().numpy() for r in range(world)]
    allrec = np.concatenate(parts) if parts else np.zeros((0, 3), np.float32)
    return np.ascontiguousarray(allrec[:, 0]).view(np.uint32), np.ascontiguousarray(allrec[:, 1:])


class ExchangePlan:
    """Halo-exchange plan for a decomposition that is kept for many frames (like the persistent communication plans of
    MPI halo exchanges).  Built ONCE from the ids each rank exports and the ids it holds as halo copies: one all-gather
    of the export lists tells every rank where, in the concatenation of all ranks' exports, its ghosts' moves will
    arrive.  Per frame `exchange(dxy)` is then one fixed-shape all-gather of the 8-byte moves (NCCL over NVLink on
    GPUs, gloo in the CPU tests) and one gather-by-index; no ids travel, nothing is searched."""

    def __init__(self, export_ids: np.ndarray, ghost_ids: np.ndarray):
        self.export_ids = np.ascontiguousarray(export_ids, np.uint32)
        self.ghost_ids = np.ascontiguousarray(ghost_ids, np.uint32)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if self.world == 1:
            self.width, self.take, self.found = len(self.export_ids), np.zeros(0, np.int64), np.zeros(0, bool)
            return
        all_ids, _ = exchange_boundary_moves(self.export_ids, np.zeros((len(self.export_ids), 2), np.float32))
        dev = _device()
        counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
        counts[self.rank] = len(self.export_ids)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        host_counts = np.asarray(counts.cpu().tolist(), np.int64)
        self.width = int(max(host_counts.max(), 1))
        # position of record k of rank r in the padded (world, width) receive buffer
        starts = np.concatenate([[0], np.cumsum(host_counts)[:-1]])
        owner = np.repeat(np.arange(self.world), host_counts)
        padded_pos = owner * self.width + (np.arange(len(all_ids)) - starts[owner])
        order = np.argsort(all_ids, kind="stable")
        sorted_ids = all_ids[order]
        at = np.searchsorted(sorted_ids, self.ghost_ids)
        at = np.minimum(at, max(len(sorted_ids) - 1, 0))
        self.found = sorted_ids[at] == self.ghost_ids if len(sorted_ids) else np.zeros(len(self.ghost_ids), bool)
        self.take = padded_pos[order[at]] if len(sorted_ids) else np.zeros(len(self.ghost_ids), np.int64)
        self._send = torch.zeros((self.width, 2), dtype=torch.float32, device=dev)
        self._recv = torch.empty((self.world * self.width, 2), dtype=torch.float32, device=dev)
        pin = dev.type == "cuda"
        self._host_send = torch.zeros((self.width, 2), dtype=torch.float32, pin_memory=pin)
        self._host_recv = torch.empty((self.world * self.width, 2), dtype=torch.float32, pin_memory=pin)

    def exchange(self, dxy: np.ndarray) -> np.ndarray:
        """dxy: this rank's moves of `export_ids` (same order).  Returns the moves of `ghost_ids` (same order; rows of
        ghosts nobody exported — see `found` — are zero)."""
        dxy = np.ascontiguousarray(dxy, np.float32).reshape(-1, 2)
        if self.world == 1:
            return np.zeros((len(self.ghost_ids), 2), np.float32)
        n = len(self.export_ids)
        self._host_send.numpy()[:n] = dxy
        self._send.copy_(self._host_send, non_blocking=True)
        dist.all_gather_into_tensor(self._recv, self._send)
        self._host_recv.copy_(self._recv, non_blocking=False)
        out = self._host_recv.numpy()[self.take]
        out[~self.found] = 0.0
        return out
