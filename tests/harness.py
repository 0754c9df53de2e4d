"""Backend-agnostic test harness over the flat ``rn_*`` interface (tests/cpp/runner_api.h).

TEST INFRASTRUCTURE.  Three shared libraries export the same symbols:

* ``oracle/libc2d_oracle.so``            plain-C restatement of the reference (the oracle)
* ``oracle/_ref/libc2d_ref_runner.so``   the UNMODIFIED reference, compiled here from /root/reference
* ``colli2de_b200/libc2d_b200_runner.so`` this repo's ``c2d::Registry`` (CUDA through the C-ABI)

so one scripted scene can be replayed on any of them and the results compared.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LIB_PATHS = {
    "oracle": os.path.join(REPO, "oracle", "libc2d_oracle.so"),
    "reference": os.path.join(REPO, "oracle", "_ref", "libc2d_ref_runner.so"),
    "b200": os.environ.get("C2D_B200_RUNNER", os.path.join(REPO, "colli2de_b200", "libc2d_b200_runner.so")),
}

PAIR_DTYPE = np.dtype(
    [
        ("entityA", "<u4"),
        ("entityB", "<u4"),
        ("shapeA", "<u4"),
        ("shapeB", "<u4"),
        ("normal", "<f4", (2,)),
        ("points", [("point", "<f4", (2,)), ("anchorA", "<f4", (2,)), ("anchorB", "<f4", (2,)), ("separation", "<f4")], (2,)),
        ("pointCount", "u1"),
        ("_pad", "u1", (3,)),
    ]
)
assert PAIR_DTYPE.itemsize == 84

MANIFOLD_DTYPE = np.dtype(
    [
        ("normal", "<f4", (2,)),
        ("points", [("point", "<f4", (2,)), ("anchorA", "<f4", (2,)), ("anchorB", "<f4", (2,)), ("separation", "<f4")], (2,)),
        ("pointCount", "u1"),
        ("_pad", "u1", (3,)),
    ]
)
assert MANIFOLD_DTYPE.itemsize == 68

HIT_DTYPE = np.dtype(
    [
        ("entity", "<u4"),
        ("shape", "<u4"),
        ("entry", "<f4", (2,)),
        ("exit", "<f4", (2,)),
        ("entryTime", "<f4"),
        ("exitTime", "<f4"),
    ]
)
assert HIT_DTYPE.itemsize == 32

STATIC, DYNAMIC, BULLET = 0, 1, 2
CIRCLE, CAPSULE, SEGMENT, POLYGON = 0, 1, 2, 3


def _p(a, ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


_LIBS: dict[str, C.CDLL] = {}


def available(name: str) -> bool:
    return os.path.exists(LIB_PATHS[name])


def load(name: str) -> C.CDLL:
    if name in _LIBS:
        return _LIBS[name]
    path = LIB_PATHS[name]
    # RTLD_LOCAL: the reference runner and the B200 runner both instantiate c2d::Registry<uint32_t> (same mangled
    # names, different code) and must never see each other's symbols; both are also linked -Bsymbolic.
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    pf, pu8, pu32, pu64 = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    pd = C.POINTER(C.c_double)
    sig = {
        "rn_backend_name": (C.c_char_p, []),
        "rn_create": (vp, [i32, u32]),
        "rn_destroy": (None, [vp]),
        "rn_last_error": (C.c_char_p, [vp]),
        "rn_create_entities": (None, [vp, u32, pu32, pu8, pf]),
        "rn_create_entities_raw": (None, [vp, u32, pu32, pu8, pf]),
        "rn_add_shapes": (None, [vp, u32, pu32, pu8, pf, pu8, pu64, pu64, pu32]),
        "rn_remove_shapes": (None, [vp, u32, pu32]),
        "rn_remove_entities": (None, [vp, u32, pu32]),
        "rn_set_active": (None, [vp, u32, pu32, pu8]),
        "rn_move_translate": (None, [vp, u32, pu32, pf]),
        "rn_move_transform": (None, [vp, u32, pu32, pf]),
        "rn_move_entities_bulk": (None, [vp, u32, pu32, pf, i32, pd]),
        "rn_teleport_transform": (None, [vp, u32, pu32, pf]),
        "rn_teleport_translation": (None, [vp, u32, pu32, pf]),
        "rn_get_transforms": (None, [vp, u32, pu32, pf]),
        "rn_get_prev_transforms": (None, [vp, u32, pu32, pf]),
        "rn_size": (u64, [vp]),
        "rn_clear": (None, [vp]),
        "rn_get_colliding_pairs": (u64, [vp, pd]),
        "rn_copy_pairs": (None, [vp, vp]),
        "rn_count_colliding_pairs_device": (u64, [vp, pd]),
        "rn_query_stats": (None, [vp, pd]),
        "rn_native_context": (vp, [vp]),
        "rn_shape_slots": (u32, [vp]),
        "rn_read_boxes": (None, [vp, u32, u32, pf, pf, pu8]),
        "rn_candidates": (u64, [vp]),
        "rn_copy_candidates": (None, [vp, pu32]),
        "rn_set_ghost": (None, [vp, u32, pu32, pu8]),
        "rn_get_colliding_pairs_view": (u64, [vp, pd, C.POINTER(vp)]),
        "rn_group_id": (None, [vp, vp]),
        "rn_group_join": (None, [vp, u32, u32, vp]),
        "rn_group_options": (None, [vp, i32, i32]),
        "rn_group_leave": (None, [vp]),
        "rn_group_ray_sharding": (None, [vp, i32]),
        "rn_get_collisions": (u64, [vp, u32]),
        "rn_get_collisions_many": (u64, [vp, u32, pu32, pd]),
        "rn_serialize": (u64, [vp]),
        "rn_copy_blob": (None, [vp, vp]),
        "rn_deserialize": (vp, [i32, vp, u64]),
        "rn_seed_entities": (None, [vp, u32, pu32, pu8, pf]),
        "rn_equals": (i32, [vp, vp]),
        "rn_are_colliding": (None, [vp, u32, pu32, pu32, pu8]),
        "rn_first_hit_raycast": (None, [vp, u32, pf, pu8, pu64, pu8, vp]),
        "rn_stage_translations": (i32, [vp, u32, pu32, pf]),
        "rn_apply_staged": (None, [vp, i32]),
        "rn_release_staged": (None, [vp, i32]),
        "rn_run_staged_frames": (u64, [vp, u32, C.POINTER(C.c_int), pd, pd]),
        "rn_raycast": (u64, [vp, u32, pf, pu8, pu64, pu64, pd]),
        "rn_copy_hits": (None, [vp, vp]),
        "rn_collide": (None, [u32, pu8, pf, pu8, pf, pu8, pf, pu8, pf, vp, pu8]),
        "rn_sweep": (None, [u32, i32, pu8, pf, pu8, pf, pf, pu8, pf, pu8, pf, pf, pf, vp, pu8]),
        "rn_raycast_shape": (None, [u32, pu8, pf, pu8, pf, pf, pu8, pf, pu8]),
        "rn_sweep_ray": (None, [u32, pu8, pf, pu8, pf, pf, pf, pu8, pf, pu8]),
        "rn_compute_aabb": (None, [u32, pu8, pf, pu8, pf, pf]),
        "rn_polygon_normals": (None, [u32, pf, pu8, pf]),
    }
    optional = {"rn_serialize", "rn_copy_blob", "rn_deserialize", "rn_seed_entities", "rn_equals", "rn_group_id",
                "rn_group_join", "rn_group_options", "rn_group_leave", "rn_group_ray_sharding", "rn_run_staged_frames", "rn_move_entities_bulk"}  # not in the C oracle
    for fn, (res, args) in sig.items():
        if fn in optional and not hasattr(lib, fn):
            continue
        f = getattr(lib, fn)
        f.restype = res
        f.argtypes = args
    _LIBS[name] = lib
    return lib


def xf5_from_xf3(xf3) -> np.ndarray:
    """Transform(Vec2, angle): sin/cos from libm's float sinf/cosf (Transform.hpp:23-28)."""
    xf3 = _f32(xf3).reshape(-1, 3)
    out = np.zeros((xf3.shape[0], 5), np.float32)
    out[:, :3] = xf3
    libm = C.CDLL("libm.so.6")
    libm.sinf.restype = C.c_float
    libm.sinf.argtypes = [C.c_float]
    libm.cosf.restype = C.c_float
    libm.cosf.argtypes = [C.c_float]
    for i in range(xf3.shape[0]):
        out[i, 3] = libm.sinf(float(xf3[i, 2]))
        out[i, 4] = libm.cosf(float(xf3[i, 2]))
    return out


class Backend:
    """A Registry<uint32_t, Method> living behind one of the rn_* libraries."""

    def __init__(self, name: str, method: int = 0, cell_size: int = 120, _ctx=None):
        self.name = name
        self.method = method
        self.lib = load(name)
        self.ctx = _ctx if _ctx is not None else self.lib.rn_create(method, cell_size)
        self._check()

    # -- snapshots ----------------------------------------------------------------------------
    def serialize(self) -> bytes:
        """Registry::serialize -> the snapshot bytes."""
        n = int(self.lib.rn_serialize(self.ctx))
        self._check()
        buf = C.create_string_buffer(n)
        self.lib.rn_copy_blob(self.ctx, buf)
        return buf.raw

    @classmethod
    def deserialize(cls, name: str, blob: bytes, method: int = 0, entities=None):
        """Registry::deserialize(blob) in backend `name`.  `entities` = (ids, body, prev_xf5) re-seeds the harness's own
        mirror of body types / bullets' previous transforms (it is kept by observation, not read from the Registry)."""
        lib = load(name)
        ctx = lib.rn_deserialize(method, blob, len(blob))
        b = cls(name, method, _ctx=ctx)
        if entities is not None:
            ids, body, prev = _u32(entities[0]), _u8(entities[1]), _f32(entities[2])
            lib.rn_seed_entities(b.ctx, len(ids), _p(ids, C.c_uint32), _p(body, C.c_uint8), _p(prev, C.c_float))
            b._check()
        return b

    def equals(self, other) -> bool:
        r = int(self.lib.rn_equals(self.ctx, other.ctx))
        self._check()
        return r == 1

    def close(self):
        if self.ctx:
            self.lib.rn_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self):
        err = self.lib.rn_last_error(self.ctx)
        if err:
            raise RuntimeError(f"[{self.name}] {err.decode()}")

    # -- mutation -----------------------------------------------------------------------------
    def create_entities(self, ids, body, xf3):
        ids, body, xf3 = _u32(ids), _u8(body), _f32(xf3)
        self.lib.rn_create_entities(self.ctx, len(ids), _p(ids, C.c_uint32), _p(body, C.c_uint8), _p(xf3, C.c_float))
        self._check()

    def create_entities_raw(self, ids, body, xf5):
        ids, body, xf5 = _u32(ids), _u8(body), _f32(xf5)
        self.lib.rn_create_entities_raw(self.ctx, len(ids), _p(ids, C.c_uint32), _p(body, C.c_uint8), _p(xf5, C.c_float))
        self._check()

    def add_shapes(self, entity, stype, geom16, vcount=None, cat=None, mask=None) -> np.ndarray:
        entity, stype, geom16 = _u32(entity), _u8(stype), _f32(geom16)
        n = len(entity)
        assert geom16.shape == (n, 16)
        vcount = _u8(vcount if vcount is not None else np.zeros(n))
        cat = _u64(cat) if cat is not None else None
        mask = _u64(mask) if mask is not None else None
        out = np.zeros(n, np.uint32)
        self.lib.rn_add_shapes(
            self.ctx, n, _p(entity, C.c_uint32), _p(stype, C.c_uint8), _p(geom16, C.c_float), _p(vcount, C.c_uint8),
            _p(cat, C.c_uint64), _p(mask, C.c_uint64), _p(out, C.c_uint32))
        self._check()
        return out

    def remove_shapes(self, shape_ids):
        s = _u32(shape_ids)
        self.lib.rn_remove_shapes(self.ctx, len(s), _p(s, C.c_uint32))
        self._check()

    def remove_entities(self, ids):
        s = _u32(ids)
        self.lib.rn_remove_entities(self.ctx, len(s), _p(s, C.c_uint32))
        self._check()

    def set_active(self, shape_ids, active):
        s, a = _u32(shape_ids), _u8(active)
        self.lib.rn_set_active(self.ctx, len(s), _p(s, C.c_uint32), _p(a, C.c_uint8))
        self._check()

    def set_ghost(self, shape_ids, ghost):
        s, g = _u32(shape_ids), _u8(ghost)
        self.lib.rn_set_ghost(self.ctx, len(s), _p(s, C.c_uint32), _p(g, C.c_uint8))
        self._check()

    # multi-GPU group of Registries (B200 only): Registry::createGroupId / joinGroup / setGroupResults + setGroupMoveExchange
    def group_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self.lib.rn_group_id(self.ctx, buf)
        self._check()
        return buf.raw

    def group_join(self, rank: int, world: int, ident: bytes):
        self.lib.rn_group_join(self.ctx, rank, world, C.c_char_p(ident))
        self._check()

    def group_join_torch(self):
        """Joins a group spanning the torch.distributed world (the id travels by broadcast_object_list)."""
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        ident = [self.group_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        self.group_join(rank, world, ident[0])

    def group_options(self, results: int = 0, exchange: bool = False):
        self.lib.rn_group_options(self.ctx, results, 1 if exchange else 0)
        self._check()

    def group_ray_sharding(self, on: bool):
        self.lib.rn_group_ray_sharding(self.ctx, 1 if on else 0)
        self._check()

    def group_leave(self):
        self.lib.rn_group_leave(self.ctx)
        self._check()

    def move_translate(self, ids, dxy):
        ids, dxy = _u32(ids), _f32(dxy)
        self.lib.rn_move_translate(self.ctx, len(ids), _p(ids, C.c_uint32), _p(dxy, C.c_float))
        self._check()

    def move_entities_bulk(self, ids, dxy, teleport: bool = False) -> float:
        """Registry::moveEntities / teleportEntities(positions) in one call (reference: the loop); returns its seconds"""
        ids, dxy = _u32(ids), _f32(dxy)
        sec = C.c_double(0.0)
        self.lib.rn_move_entities_bulk(self.ctx, len(ids), _p(ids, C.c_uint32), _p(dxy, C.c_float), 1 if teleport else 0, C.byref(sec))
        self._check()
        return sec.value

    def move_transform(self, ids, dxf3):
        ids, d = _u32(ids), _f32(dxf3)
        self.lib.rn_move_transform(self.ctx, len(ids), _p(ids, C.c_uint32), _p(d, C.c_float))
        self._check()

    def teleport_transform(self, ids, xf3):
        ids, d = _u32(ids), _f32(xf3)
        self.lib.rn_teleport_transform(self.ctx, len(ids), _p(ids, C.c_uint32), _p(d, C.c_float))
        self._check()

    def teleport_translation(self, ids, xy):
        ids, d = _u32(ids), _f32(xy)
        self.lib.rn_teleport_translation(self.ctx, len(ids), _p(ids, C.c_uint32), _p(d, C.c_float))
        self._check()

    def get_transforms(self, ids) -> np.ndarray:
        ids = _u32(ids)
        out = np.zeros((len(ids), 5), np.float32)
        self.lib.rn_get_transforms(self.ctx, len(ids), _p(ids, C.c_uint32), _p(out, C.c_float))
        self._check()
        return out

    def get_prev_transforms(self, ids) -> np.ndarray:
        ids = _u32(ids)
        out = np.zeros((len(ids), 5), np.float32)
        self.lib.rn_get_prev_transforms(self.ctx, len(ids), _p(ids, C.c_uint32), _p(out, C.c_float))
        self._check()
        return out

    def size(self) -> int:
        return int(self.lib.rn_size(self.ctx))

    def clear(self):
        self.lib.rn_clear(self.ctx)
        self._check()

    # -- queries ------------------------------------------------------------------------------
    def get_colliding_pairs(self, with_time=False):
        sec = C.c_double(0.0)
        n = int(self.lib.rn_get_colliding_pairs(self.ctx, C.byref(sec)))
        self._check()
        out = np.zeros(n, PAIR_DTYPE)
        if n:
            self.lib.rn_copy_pairs(self.ctx, out.ctypes.data_as(C.c_void_p))
        return (out, sec.value) if with_time else out

    STAT_FIELDS = ("shapes_alive", "candidates", "pairs", "h2d_bytes", "d2h_bytes", "kernel_launches", "retries",
                   "ms_refit", "ms_build", "ms_broad", "ms_narrow", "ms_device", "ms_total", "ms_host_enqueue", "ms_host_wait")

    def count_pairs_device(self):
        sec = C.c_double(0.0)
        n = int(self.lib.rn_count_colliding_pairs_device(self.ctx, C.byref(sec)))
        self._check()
        return n, sec.value

    def native_context(self):
        return self.lib.rn_native_context(self.ctx)

    # -- secondary oracles ----------------------------------------
    def shape_slots(self) -> int:
        n = int(self.lib.rn_shape_slots(self.ctx))
        self._check()
        return n

    def read_boxes(self, first: int = 0, count: int | None = None):
        """(tight (n,4), leaf (n,4), alive (n,)): ShapeInstance::mBoundingBox, the shape's leaf box in its tree, liveness."""
        if count is None:
            count = self.shape_slots() - first
        tight, fat = np.zeros((count, 4), np.float32), np.zeros((count, 4), np.float32)
        alive = np.zeros(count, np.uint8)
        self.lib.rn_read_boxes(self.ctx, first, count, _p(tight, C.c_float), _p(fat, C.c_float), _p(alive, C.c_uint8))
        self._check()
        return tight, fat, alive

    def candidates(self) -> np.ndarray:
        """Broadphase candidate pairs (shapeA, shapeB) of the five groups; same-tree pairs as (min, max)."""
        n = int(self.lib.rn_candidates(self.ctx))
        self._check()
        out = np.zeros((n, 2), np.uint32)
        if n:
            self.lib.rn_copy_candidates(self.ctx, _p(out, C.c_uint32))
        return out

    def stats(self) -> dict:
        out = np.zeros(16, np.float64)
        self.lib.rn_query_stats(self.ctx, _p(out, C.c_double))
        return dict(zip(self.STAT_FIELDS, out.tolist()))

    def run_staged_frames(self, handles):
        """apply_staged(h) + count_pairs_device() for every handle in native code: (summed pairs, summed stats, seconds)"""
        hs = (C.c_int * len(handles))(*handles)
        sums = np.zeros(16, np.float64)
        sec = C.c_double(0.0)
        pairs = int(self.lib.rn_run_staged_frames(self.ctx, len(handles), hs, _p(sums, C.c_double), C.byref(sec)))
        self._check()
        return pairs, dict(zip(self.STAT_FIELDS, sums.tolist())), sec.value

    def stage_translations(self, ids, dxy) -> int:
        ids, dxy = _u32(ids), _f32(dxy)
        h = int(self.lib.rn_stage_translations(self.ctx, len(ids), _p(ids, C.c_uint32), _p(dxy, C.c_float)))
        self._check()
        return h

    def apply_staged(self, handle: int):
        self.lib.rn_apply_staged(self.ctx, handle)
        self._check()

    def release_staged(self, handle: int):
        self.lib.rn_release_staged(self.ctx, handle)
        self._check()

    def get_colliding_pairs_view(self):
        """(records copied out of the view, seconds of the API call only)"""
        sec, ptr = C.c_double(0.0), C.c_void_p()
        n = int(self.lib.rn_get_colliding_pairs_view(self.ctx, C.byref(sec), C.byref(ptr)))
        self._check()
        out = np.zeros(n, PAIR_DTYPE)
        if n:
            C.memmove(out.ctypes.data, ptr.value, n * PAIR_DTYPE.itemsize)
        return out, sec.value

    def get_collisions(self, entity_id: int):
        n = int(self.lib.rn_get_collisions(self.ctx, int(entity_id)))
        self._check()
        out = np.zeros(n, PAIR_DTYPE)
        if n:
            self.lib.rn_copy_pairs(self.ctx, out.ctypes.data_as(C.c_void_p))
        return out

    def get_collisions_many(self, entity_ids):
        """getCollisions() looped natively over the ids; returns (summed result size, seconds)."""
        ids = _u32(entity_ids)
        sec = C.c_double(0.0)
        total = int(self.lib.rn_get_collisions_many(self.ctx, len(ids), _p(ids, C.c_uint32), C.byref(sec)))
        self._check()
        return total, sec.value

    def are_colliding(self, ids_a, ids_b) -> np.ndarray:
        a, b = _u32(ids_a), _u32(ids_b)
        out = np.zeros(len(a), np.uint8)
        self.lib.rn_are_colliding(self.ctx, len(a), _p(a, C.c_uint32), _p(b, C.c_uint32), _p(out, C.c_uint8))
        self._check()
        return out

    def first_hit_raycast(self, ray4, infinite=None, masks=None):
        ray4 = _f32(ray4).reshape(-1, 4)
        n = ray4.shape[0]
        infinite = _u8(infinite if infinite is not None else np.zeros(n))
        masks = _u64(masks) if masks is not None else None
        found = np.zeros(n, np.uint8)
        hits = np.zeros(n, HIT_DTYPE)
        self.lib.rn_first_hit_raycast(self.ctx, n, _p(ray4, C.c_float), _p(infinite, C.c_uint8), _p(masks, C.c_uint64),
                                      _p(found, C.c_uint8), hits.ctypes.data_as(C.c_void_p))
        self._check()
        return found, hits

    def raycast(self, ray4, infinite=None, masks=None, with_time=False):
        ray4 = _f32(ray4).reshape(-1, 4)
        n = ray4.shape[0]
        infinite = _u8(infinite if infinite is not None else np.zeros(n))
        masks = _u64(masks) if masks is not None else None
        offsets = np.zeros(n + 1, np.uint64)
        sec = C.c_double(0.0)
        total = int(self.lib.rn_raycast(self.ctx, n, _p(ray4, C.c_float), _p(infinite, C.c_uint8),
                                        _p(masks, C.c_uint64), _p(offsets, C.c_uint64), C.byref(sec)))
        self._check()
        hits = np.zeros(total, HIT_DTYPE)
        if total:
            self.lib.rn_copy_hits(self.ctx, hits.ctypes.data_as(C.c_void_p))
        return (offsets, hits, sec.value) if with_time else (offsets, hits)


# -- free functions (batched) ---------------------------------------------------------------------
@dataclass
class ShapeBatch:
    stype: np.ndarray   # (n,) u8
    geom: np.ndarray    # (n,16) f32
    count: np.ndarray   # (n,) u8

    def __post_init__(self):
        self.stype = _u8(self.stype)
        self.geom = _f32(self.geom).reshape(-1, 16)
        self.count = _u8(self.count)

    def __len__(self):
        return len(self.stype)

    def take(self, idx):
        return ShapeBatch(self.stype[idx], self.geom[idx], self.count[idx])


def collide(name, A: ShapeBatch, xfA5, B: ShapeBatch, xfB5):
    lib = load(name)
    n = len(A)
    xfA5, xfB5 = _f32(xfA5).reshape(n, 5), _f32(xfB5).reshape(n, 5)
    man = np.zeros(n, MANIFOLD_DTYPE)
    flag = np.zeros(n, np.uint8)
    lib.rn_collide(n, _p(A.stype, C.c_uint8), _p(A.geom, C.c_float), _p(A.count, C.c_uint8), _p(xfA5, C.c_float),
                   _p(B.stype, C.c_uint8), _p(B.geom, C.c_float), _p(B.count, C.c_uint8), _p(xfB5, C.c_float),
                   man.ctypes.data_as(C.c_void_p), _p(flag, C.c_uint8))
    return man, flag


def sweep(name, mode, A: ShapeBatch, xfA0, xfA1, B: ShapeBatch, xfB0, xfB1=None):
    lib = load(name)
    n = len(A)
    xfA0, xfA1, xfB0 = _f32(xfA0).reshape(n, 5), _f32(xfA1).reshape(n, 5), _f32(xfB0).reshape(n, 5)
    xfB1 = _f32(xfB1).reshape(n, 5) if xfB1 is not None else xfB0.copy()
    frac = np.zeros(n, np.float32)
    man = np.zeros(n, MANIFOLD_DTYPE)
    hit = np.zeros(n, np.uint8)
    lib.rn_sweep(n, mode, _p(A.stype, C.c_uint8), _p(A.geom, C.c_float), _p(A.count, C.c_uint8),
                 _p(xfA0, C.c_float), _p(xfA1, C.c_float), _p(B.stype, C.c_uint8), _p(B.geom, C.c_float),
                 _p(B.count, C.c_uint8), _p(xfB0, C.c_float), _p(xfB1, C.c_float), _p(frac, C.c_float),
                 man.ctypes.data_as(C.c_void_p), _p(hit, C.c_uint8))
    return frac, man, hit


def raycast_shape(name, S: ShapeBatch, xf5, ray4, infinite):
    lib = load(name)
    n = len(S)
    xf5, ray4, infinite = _f32(xf5).reshape(n, 5), _f32(ray4).reshape(n, 4), _u8(infinite)
    times = np.zeros((n, 2), np.float32)
    hit = np.zeros(n, np.uint8)
    lib.rn_raycast_shape(n, _p(S.stype, C.c_uint8), _p(S.geom, C.c_float), _p(S.count, C.c_uint8),
                         _p(xf5, C.c_float), _p(ray4, C.c_float), _p(infinite, C.c_uint8), _p(times, C.c_float),
                         _p(hit, C.c_uint8))
    return times, hit


def sweep_ray(name, S: ShapeBatch, xf5_0, xf5_1, ray4, infinite):
    lib = load(name)
    n = len(S)
    a, b = _f32(xf5_0).reshape(n, 5), _f32(xf5_1).reshape(n, 5)
    ray4, infinite = _f32(ray4).reshape(n, 4), _u8(infinite)
    times = np.zeros((n, 2), np.float32)
    hit = np.zeros(n, np.uint8)
    lib.rn_sweep_ray(n, _p(S.stype, C.c_uint8), _p(S.geom, C.c_float), _p(S.count, C.c_uint8), _p(a, C.c_float),
                     _p(b, C.c_float), _p(ray4, C.c_float), _p(infinite, C.c_uint8), _p(times, C.c_float),
                     _p(hit, C.c_uint8))
    return times, hit


def compute_aabb(name, S: ShapeBatch, xf5):
    lib = load(name)
    n = len(S)
    xf5 = _f32(xf5).reshape(n, 5)
    out = np.zeros((n, 4), np.float32)
    lib.rn_compute_aabb(n, _p(S.stype, C.c_uint8), _p(S.geom, C.c_float), _p(S.count, C.c_uint8),
                        _p(xf5, C.c_float), _p(out, C.c_float))
    return out


def polygon_normals(name, geom, count):
    lib = load(name)
    geom, count = _f32(geom).reshape(-1, 16), _u8(count)
    out = np.zeros((len(count), 16), np.float32)
    lib.rn_polygon_normals(len(count), _p(geom, C.c_float), _p(count, C.c_uint8), _p(out, C.c_float))
    return out
