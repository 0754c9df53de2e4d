"""N3: wire compatibility of Registry snapshots, host side.

A snapshot written by the REFERENCE (oracle/_ref) is decoded by this repo's codec
(include/colli2de/internal/utils/Snapshot.hpp, through tests/cpp/snapshot_check.cpp), re-encoded — which rebuilds the
three DynamicBVH images from the leaf boxes — and loaded back INTO THE REFERENCE.  The reloaded registry must answer
getCollidingPairs / rayCast exactly like the original, also after both were mutated further (the re-built trees have
to be valid trees the reference can insert into, move in and remove from; the tight and leaf boxes are history and
must have travelled bit for bit).
"""
import ctypes as C
import os

import numpy as np
import pytest

from colli2de_b200 import scenes
from tests import harness

SNAP_LIB = os.path.join(harness.REPO, "colli2de_b200", "libc2d_snapshot_check.so")
needs = pytest.mark.skipif(not (harness.available("reference") and os.path.exists(SNAP_LIB)),
                           reason="reference runner / snapshot_check not built")


def snap_lib():
    lib = C.CDLL(SNAP_LIB)
    lib.snap_last_error.restype = C.c_char_p
    lib.snap_reencode.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    lib.snap_summary.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    lib.snap_free.argtypes = [C.c_void_p]
    return lib


def reencode(blob: bytes, grid: bool = False) -> bytes:
    lib = snap_lib()
    out, n = C.c_void_p(), C.c_uint64()
    rc = lib.snap_reencode(blob, len(blob), int(grid), C.byref(out), C.byref(n))
    assert rc == 0, lib.snap_last_error().decode()
    data = C.string_at(out, n.value)
    lib.snap_free(out)
    return data


def summary(blob: bytes, grid: bool = False):
    lib = snap_lib()
    counts = (C.c_uint64 * 6)()
    assert lib.snap_summary(blob, len(blob), int(grid), counts) == 0, lib.snap_last_error().decode()
    return list(counts)


def build_history(b, seed=5):
    """A registry with history: mixed shapes of all body types, moves (leaf boxes grow), removals (free ids, stale
    slots), inactive shapes, rotated bullets."""
    sc = scenes.mixed_scene(300, 60, 90.0, seed=seed)
    sc["ent_body"][::17] = 2  # some bullets
    shape_ids = scenes.populate(b, sc)
    rng = np.random.default_rng(seed)
    ids = sc["ent_id"]
    movable = ids[sc["ent_body"] != 0]
    for _ in range(3):
        b.move_translate(movable, rng.uniform(-1.5, 1.5, (len(movable), 2)).astype(np.float32))
    d3 = np.zeros((len(movable), 3), np.float32)
    d3[:, :2] = rng.uniform(-0.5, 0.5, (len(movable), 2))
    d3[:, 2] = rng.uniform(-0.3, 0.3, len(movable))
    b.move_transform(movable, d3)
    gone_entities = ids[5::23]
    b.remove_entities(gone_entities)
    alive_shapes = shape_ids[~np.isin(sc["shp_entity"], gone_entities)]
    b.remove_shapes(alive_shapes[3::29])
    alive_shapes = np.setdiff1d(alive_shapes, alive_shapes[3::29])
    b.set_active(alive_shapes[1::19], np.zeros(len(alive_shapes[1::19]), np.uint8))
    keep = ~np.isin(ids, gone_entities)
    return sc, ids[keep], sc["ent_body"][keep], shape_ids


def pair_table(b):
    recs = b.get_colliding_pairs()
    order = np.lexsort((recs["shapeB"], recs["shapeA"], recs["entityB"], recs["entityA"]))
    return recs[order]


def same_pairs(a, b, unique=True):
    """Equal as SETS of shape pairs; records that come out in the same A/B orientation must be byte-identical (for
    same-tree pairs the reference's orientation depends on its tree topology, SURVEY.md D8, and the re-encoded snapshot
    carries rebuilt trees)."""
    def keyed(recs):
        recs = recs.copy()
        recs["_pad"] = 0  # three bytes of struct padding after Manifold::pointCount: indeterminate in the reference
        return {(min(int(r["shapeA"]), int(r["shapeB"])), max(int(r["shapeA"]), int(r["shapeB"]))): r for r in recs}
    ka, kb = keyed(a), keyed(b)
    # PartitioningMethod::Grid reports a pair once per shared cell: only the SET is comparable there
    if ka.keys() != kb.keys() or (unique and (len(ka) != len(a) or len(kb) != len(b))):
        return False
    same_orientation = 0
    for k, ra in ka.items():
        rb = kb[k]
        if ra["shapeA"] == rb["shapeA"]:
            same_orientation += 1
            if ra.tobytes() != rb.tobytes():
                return False
    return same_orientation > 0 or len(ka) == 0


@needs
@pytest.mark.parametrize("method", [0, 1], ids=["none", "grid"])
def test_reference_snapshot_roundtrip_through_codec(method):
    grid = method == 1
    ref = harness.Backend("reference", method, 16)
    sc, ids, body, _ = build_history(ref)
    blob = ref.serialize()
    counts = summary(blob, grid)
    assert counts[0] == len(ids) and counts[3] > 0 and counts[5] > 0 and counts[4] == int((body == 2).sum())

    again = reencode(blob, grid)
    assert summary(again, grid) == counts
    prev = ref.get_prev_transforms(ids)
    clone = harness.Backend.deserialize("reference", again, method=method, entities=(ids, body, prev))
    assert clone.size() == ref.size()

    a, b = pair_table(ref), pair_table(clone)
    assert len(a) > 0 and same_pairs(a, b, unique=not grid)

    # both registries keep living: the rebuilt trees take moves, removals and insertions like the originals
    rng = np.random.default_rng(11)
    movable = ids[body != 0]
    for step in range(3):
        d = rng.uniform(-2.0, 2.0, (len(movable), 2)).astype(np.float32)
        for reg in (ref, clone):
            reg.move_translate(movable, d)
        if step == 1:
            for reg in (ref, clone):
                reg.remove_entities(ids[7::31])
                reg.create_entities(np.uint32([900001, 900002]), np.uint8([1, 2]), np.float32([[10, 10, 0.2], [40, 40, 0.0]]))
                g = np.zeros((2, 16), np.float32)
                g[:, 2] = 3.0
                got = reg.add_shapes(np.uint32([900001, 900002]), np.uint8([0, 0]), g)
                assert len(got) == 2
            movable = np.setdiff1d(movable, ids[7::31])
        a, b = pair_table(ref), pair_table(clone)
        assert same_pairs(a, b, unique=not grid), f"step {step}"

    ray4, inf = scenes.rays(64, 90.0, seed=3, finite_len=30.0)
    oa, ha = ref.raycast(ray4, inf)
    ob, hb = clone.raycast(ray4, inf)
    if not grid:
        assert np.array_equal(oa, ob) and ha.tobytes() == hb.tobytes()
    else:
        # grid rays walk cells and test the per-cell leaf boxes, which a re-encoded snapshot re-creates from one box
        # per shape: the narrow hits agree except where leaf-box (entry, exit) ties resolve differently
        same = sum(1 for i in range(64) if ha[int(oa[i]):int(oa[i + 1])].tobytes() == hb[int(ob[i]):int(ob[i + 1])].tobytes())
        assert same >= 60, same


@needs
def test_reencoded_snapshot_of_reference_test_scene():
    """The scene of the reference's own serialisation test (tests/registry/test_registry_basic.cpp:386-409): 1000
    dynamic entities, circle / rectangle / hexagon, three category-mask combinations."""
    ref = harness.Backend("reference")
    n = 1000
    i = np.arange(n)
    xf3 = np.zeros((n, 3), np.float32)
    xf3[:, 0] = i.astype(np.float32)
    xf3[:, 1] = (i * np.float32(3.23)).astype(np.float32)
    ref.create_entities(i.astype(np.uint32), np.ones(n, np.uint8), xf3)
    stype = np.where(i % 3 == 0, 0, 3).astype(np.uint8)
    geom = np.zeros((n, 16), np.float32)
    count = np.zeros(n, np.uint8)
    circ = i % 3 == 0
    geom[circ, 0] = (i[circ] * np.float32(5.23)).astype(np.float32)
    geom[circ, 1] = (i[circ] * np.float32(-0.973)).astype(np.float32)
    geom[circ, 2] = 1.0
    rect = i % 3 == 1
    geom[rect, :8] = np.float32([-1, -1, 1, -1, 1, 1, -1, 1])  # makeRectangle({0,0}, 2, 2)
    count[rect] = 4
    hexa = i % 3 == 2
    ang = np.float32(5.0) + np.arange(6, dtype=np.float32) * np.float32(2 * np.pi / 6)
    for k in range(6):
        geom[hexa, 2 * k] = (i[hexa] * np.float32(2.43)).astype(np.float32) + np.cos(ang[k])
        geom[hexa, 2 * k + 1] = (i[hexa] * np.float32(-1.9)).astype(np.float32) + np.sin(ang[k])
    count[hexa] = 6
    cat = np.where(circ, 1, np.where(rect, 2, 3)).astype(np.uint64)
    mask = np.where(circ, 2, np.where(rect, 1, ~np.uint64(0))).astype(np.uint64)
    ref.add_shapes(i.astype(np.uint32), stype, geom, count, cat, mask)

    blob = ref.serialize()
    same = harness.Backend.deserialize("reference", blob)
    assert ref.equals(same)  # the reference's own round trip, operator== as in its test

    clone = harness.Backend.deserialize("reference", reencode(blob))
    a, b = pair_table(ref), pair_table(clone)
    assert same_pairs(a, b)
    assert clone.size() == ref.size() == n


@needs
def test_codec_rejects_truncated_and_corrupt_snapshots():
    ref = harness.Backend("reference")
    build_history(ref, seed=9)
    blob = ref.serialize()
    lib = snap_lib()
    out, n = C.c_void_p(), C.c_uint64()
    assert lib.snap_reencode(blob[: len(blob) // 2], len(blob) // 2, 0, C.byref(out), C.byref(n)) == 1
    assert b"end of stream" in lib.snap_last_error() or b"snapshot" in lib.snap_last_error()
    bad = bytearray(blob)
    bad[0:8] = (2 ** 62).to_bytes(8, "little")  # absurd entity count
    assert lib.snap_reencode(bytes(bad), len(bad), 0, C.byref(out), C.byref(n)) == 1
    assert b"implausible" in lib.snap_last_error()
