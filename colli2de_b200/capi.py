"""ctypes access to the C ABI of libc2d_gpu.so (include/c2d_gpu.h) for bench.py / tests.

Fails loudly when the library is missing: there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB = os.path.join(HERE, "libc2d_gpu.so")

EXPORTS = [
    "c2d_gpu_create", "c2d_gpu_destroy", "c2d_gpu_last_error", "c2d_gpu_clear", "c2d_gpu_set_broadphase",
    "c2d_gpu_set_shard", "c2d_gpu_mutation_epoch", "c2d_gpu_set_small_scene_limit", "c2d_gpu_set_rebuild_period", "c2d_gpu_set_profiling", "c2d_gpu_kernel_profile", "c2d_gpu_shape_add",
    "c2d_gpu_shape_remove", "c2d_gpu_shape_set_active", "c2d_gpu_shape_set_ghost", "c2d_gpu_shape_translate", "c2d_gpu_shapes_translate",
    "c2d_gpu_shape_move", "c2d_gpu_stage_translations", "c2d_gpu_apply_staged", "c2d_gpu_release_staged",
    "c2d_gpu_flush", "c2d_gpu_collide", "c2d_gpu_collide_device", "c2d_gpu_collide_begin", "c2d_gpu_prefault", "c2d_gpu_collide_end",
    "c2d_gpu_collide_fetch", "c2d_gpu_device_records", "c2d_gpu_raycast", "c2d_gpu_raycast_first", "c2d_gpu_raycast_begin", "c2d_gpu_raycast_fetch",
    "c2d_gpu_query_shapes", "c2d_gpu_pairs_colliding", "c2d_gpu_read_boxes", "c2d_gpu_read_shapes", "c2d_gpu_write_boxes",
    "c2d_gpu_read_candidates", "c2d_gpu_shape_slots", "c2d_gpu_collide_batch", "c2d_gpu_sweep_batch", "c2d_gpu_raycast_shape_batch",
    "c2d_gpu_compute_aabb_batch",
    "c2d_gpu_set_slab_mode", "c2d_gpu_comm_unique_id", "c2d_gpu_comm_init", "c2d_gpu_comm_destroy",
    "c2d_gpu_comm_set_move_exchange", "c2d_gpu_comm_set_result_mode", "c2d_gpu_comm_get_stats", "c2d_gpu_read_transforms", "c2d_gpu_set_ray_sharding", "c2d_gpu_set_profile_filter", "c2d_gpu_set_overlap", "c2d_gpu_set_incremental_refit", "c2d_gpu_set_result_order",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(GPU_LIB):
            raise RuntimeError(f"{GPU_LIB} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(colli2de_b200 has no CPU fallback)")
        _lib = C.CDLL(GPU_LIB)
        _lib.c2d_gpu_last_error.restype = C.c_char_p
        _lib.c2d_gpu_last_error.argtypes = [C.c_void_p]
        _lib.c2d_gpu_set_profiling.argtypes = [C.c_void_p, C.c_int]
        _lib.c2d_gpu_kernel_profile.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_int]
        _lib.c2d_gpu_set_shard.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        _lib.c2d_gpu_read_boxes.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.c2d_gpu_read_candidates.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64)]
    return _lib


class Stats(C.Structure):
    """c2d_gpu_stats"""
    _fields_ = [("shapes_alive", C.c_uint64), ("candidates", C.c_uint64), ("pairs", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("kernel_launches", C.c_uint32), ("retries", C.c_uint32), ("ms_refit", C.c_float),
                ("ms_build", C.c_float), ("ms_broad", C.c_float), ("ms_narrow", C.c_float), ("ms_device", C.c_float),
                ("ms_total", C.c_float), ("ms_host_enqueue", C.c_float), ("ms_host_wait", C.c_float), ("reserved", C.c_float)]


def _check(ctx, rc: int):
    if rc != 0:
        raise RuntimeError(lib().c2d_gpu_last_error(ctx).decode())


def set_profiling(ctx, on: bool):
    _check(ctx, lib().c2d_gpu_set_profiling(ctx, 1 if on else 0))


def set_profile_filter(ctx, kernel_name: str | None):
    """with profiling on, time only this kernel's launches (None = all)"""
    lib().c2d_gpu_set_profile_filter.argtypes = [C.c_void_p, C.c_char_p]
    _check(ctx, lib().c2d_gpu_set_profile_filter(ctx, kernel_name.encode() if kernel_name else None))


def set_result_order(ctx, reference_order: bool):
    """records come back in the reference's order (five groups, each sorted by (shapeA, shapeB)), byte-identical run to run"""
    lib().c2d_gpu_set_result_order.argtypes = [C.c_void_p, C.c_int]
    _check(ctx, lib().c2d_gpu_set_result_order(ctx, 1 if reference_order else 0))


def set_incremental_refit(ctx, on: bool):
    lib().c2d_gpu_set_incremental_refit.argtypes = [C.c_void_p, C.c_int]
    _check(ctx, lib().c2d_gpu_set_incremental_refit(ctx, 1 if on else 0))


def set_overlap(ctx, on: bool):
    lib().c2d_gpu_set_overlap.argtypes = [C.c_void_p, C.c_int]
    _check(ctx, lib().c2d_gpu_set_overlap(ctx, 1 if on else 0))


def set_rebuild_period(ctx, frames: int):
    lib().c2d_gpu_set_rebuild_period.argtypes = [C.c_void_p, C.c_uint32]
    _check(ctx, lib().c2d_gpu_set_rebuild_period(ctx, frames))


def set_small_scene_limit(ctx, shapes: int):
    """<= `shapes` live shapes: all-pairs broadphase + one narrowphase kernel instead of the tree pipeline (0 = never)."""
    lib().c2d_gpu_set_small_scene_limit.argtypes = [C.c_void_p, C.c_uint32]
    _check(ctx, lib().c2d_gpu_set_small_scene_limit(ctx, shapes))


def set_shard(ctx, rank: int, world: int):
    _check(ctx, lib().c2d_gpu_set_shard(ctx, rank, world))


def set_slab_mode(ctx, min_shapes: int):
    """Morton-slab build of the Dynamic tree when world > 1 and it has >= min_shapes leaves (0 = never)."""
    lib().c2d_gpu_set_slab_mode.argtypes = [C.c_void_p, C.c_uint32]
    _check(ctx, lib().c2d_gpu_set_slab_mode(ctx, min_shapes))


def set_ray_sharding(ctx, on: bool):
    """Batched rayCast casts only this rank's contiguous share of the batch (the other rays return empty lists)."""
    lib().c2d_gpu_set_ray_sharding.argtypes = [C.c_void_p, C.c_int]
    _check(ctx, lib().c2d_gpu_set_ray_sharding(ctx, 1 if on else 0))


class CommStats(C.Structure):
    _fields_ = [("exchange_bytes", C.c_uint64), ("gather_bytes", C.c_uint64), ("local_pairs", C.c_uint64),
                ("slab_owned", C.c_uint64), ("slab_leaves", C.c_uint64), ("ms_exchange", C.c_float),
                ("ms_gather", C.c_float), ("ms_slab", C.c_float)]


RESULTS_LOCAL, RESULTS_ALL, RESULTS_ROOT = 0, 1, 2


def comm_unique_id() -> bytes:
    """ncclUniqueId (128 bytes) created on this rank; hand it to every rank of the group."""
    buf = C.create_string_buffer(128)
    lib().c2d_gpu_comm_unique_id.argtypes = [C.c_void_p]
    _check(None, lib().c2d_gpu_comm_unique_id(buf))
    return buf.raw


def comm_init(ctx, rank: int, world: int, unique_id: bytes):
    lib().c2d_gpu_comm_init.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_char_p]
    _check(ctx, lib().c2d_gpu_comm_init(ctx, rank, world, unique_id))


def comm_init_torch(ctx):
    """Joins the context to a group spanning the torch.distributed world: rank 0's ncclUniqueId travels by broadcast."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    ident = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    comm_init(ctx, rank, world, ident[0])
    torch.cuda.synchronize()


def comm_destroy(ctx):
    lib().c2d_gpu_comm_destroy.argtypes = [C.c_void_p]
    _check(ctx, lib().c2d_gpu_comm_destroy(ctx))


def comm_set_move_exchange(ctx, on: bool):
    lib().c2d_gpu_comm_set_move_exchange.argtypes = [C.c_void_p, C.c_int]
    _check(ctx, lib().c2d_gpu_comm_set_move_exchange(ctx, 1 if on else 0))


def comm_set_result_mode(ctx, mode: int):
    lib().c2d_gpu_comm_set_result_mode.argtypes = [C.c_void_p, C.c_int]
    _check(ctx, lib().c2d_gpu_comm_set_result_mode(ctx, mode))


def comm_stats(ctx) -> dict:
    st = CommStats()
    lib().c2d_gpu_comm_get_stats.argtypes = [C.c_void_p, C.POINTER(CommStats)]
    _check(ctx, lib().c2d_gpu_comm_get_stats(ctx, C.byref(st)))
    return {k: getattr(st, k) for k, _ in CommStats._fields_}


def kernel_profile(ctx, reset: bool = True) -> dict:
    """{kernel name: (launches, total_ms)} accumulated since the last reset."""
    buf = C.create_string_buffer(1 << 16)
    _check(ctx, lib().c2d_gpu_kernel_profile(ctx, buf, len(buf), 1 if reset else 0))
    out = {}
    for line in buf.value.decode().splitlines():
        name, count, ms = line.split("\t")
        out[name] = (int(count), float(ms))
    return out


def device_records(ctx):
    """(device pointer, count) of the last query's 84-byte records."""
    ptr, n = C.c_void_p(), C.c_uint64(0)
    lib().c2d_gpu_device_records.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    _check(ctx, lib().c2d_gpu_device_records(ctx, C.byref(ptr), C.byref(n)))
    return ptr.value or 0, int(n.value)


def read_boxes(ctx, first: int, count: int):
    import numpy as np

    tight = np.zeros((count, 4), np.float32)
    fat = np.zeros((count, 4), np.float32)
    _check(ctx, lib().c2d_gpu_read_boxes(ctx, first, count, tight.ctypes.data_as(C.c_void_p),
                                         fat.ctypes.data_as(C.c_void_p)))
    return tight, fat


def read_candidates(ctx):
    import numpy as np

    n = C.c_uint64(0)
    _check(ctx, lib().c2d_gpu_read_candidates(ctx, 0, None, C.byref(n)))
    out = np.zeros((n.value, 2), np.uint32)
    if n.value:
        _check(ctx, lib().c2d_gpu_read_candidates(ctx, n.value, out.ctypes.data_as(C.c_void_p), C.byref(n)))
    return out
