"""N3 on the GPU: c2d::Registry::serialize / deserialize of this repo speak the reference's binary
snapshot format (reference Registry.hpp:1205-1344).

  reference registry with history --serialize--> bytes --deserialize--> B200 registry
      -> same getCollidingPairs (bit-identical manifolds), also after further moves: the tight and leaf boxes, which are
         history, travelled exactly;
  B200 registry --serialize--> bytes --deserialize--> reference registry
      -> the reference accepts the rebuilt DynamicBVH images and all three registries keep agreeing while they are
         mutated;
  operator== of the B200 Registry compares logical state; PartitioningMethod::Grid snapshots (BroadPhaseTree images) too.
"""
import numpy as np
import pytest

from colli2de_b200 import scenes
from tests import harness as H, parity as P
from tests.test_snapshot_codec import build_history, pair_table, same_pairs

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not H.available("reference"), reason="reference runner not built")


@needs_ref
def test_snapshots_travel_between_reference_and_b200():
    ref = H.Backend("reference")
    sc, ids, body, shape_ids = build_history(ref)
    idx = P.SceneIndex(sc, shape_ids)
    blob = ref.serialize()

    gpu = H.Backend.deserialize("b200", blob, entities=(ids, body, ref.get_prev_transforms(ids)))
    assert gpu.size() == ref.size()
    assert np.array_equal(ref.get_transforms(ids).view(np.uint32), gpu.get_transforms(ids).view(np.uint32))
    res = P.compare_pairs("reference", ref, idx, ref.get_colliding_pairs(), gpu.get_colliding_pairs())
    assert res["expected"] > 50 and res["not_bit_identical"] == 0, res

    rng = np.random.default_rng(21)
    movable = ids[body != 0]
    for step in range(3):
        d = rng.uniform(-1.0, 1.0, (len(movable), 2)).astype(np.float32)
        for reg in (ref, gpu):
            reg.move_translate(movable, d)
        res = P.compare_pairs("reference", ref, idx, ref.get_colliding_pairs(), gpu.get_colliding_pairs())
        assert res["expected"] > 50 and res["not_bit_identical"] == 0, (step, res)

    # ShapeId recycling survived: both hand out the same ids for new shapes
    g = np.zeros((3, 16), np.float32)
    g[:, 2] = 2.0
    for reg in (ref, gpu):
        reg.create_entities(np.uint32([800001]), np.uint8([1]), np.float32([[30, 30, 0]]))
    new_ref = ref.add_shapes(np.uint32([800001] * 3), np.uint8([0, 0, 0]), g)
    new_gpu = gpu.add_shapes(np.uint32([800001] * 3), np.uint8([0, 0, 0]), g)
    assert np.array_equal(new_ref, new_gpu)

    # and back: the reference loads what the B200 registry writes
    ids2 = np.append(ids, np.uint32(800001))
    body2 = np.append(body, np.uint8(1))
    blob2 = gpu.serialize()
    ref2 = H.Backend.deserialize("reference", blob2, entities=(ids2, body2, ref.get_prev_transforms(ids2)))
    assert ref2.size() == ref.size()
    assert same_pairs(pair_table(ref), pair_table(ref2))
    for step in range(3):
        d = rng.uniform(-1.0, 1.0, (len(movable), 2)).astype(np.float32)
        for reg in (ref, ref2, gpu):
            reg.move_translate(movable, d)
        if step == 1:
            for reg in (ref, ref2, gpu):
                reg.remove_entities(movable[4::37])
            movable = np.setdiff1d(movable, movable[4::37])
        assert same_pairs(pair_table(ref), pair_table(ref2)), step
        assert len(gpu.get_colliding_pairs()) == len(pair_table(ref2)), step


@needs_ref
def test_b200_snapshot_roundtrip_and_equality():
    gpu = H.Backend("b200")
    sc, ids, body, _ = build_history(gpu, seed=13)
    blob = gpu.serialize()
    twin = H.Backend.deserialize("b200", blob)
    assert gpu.equals(twin) and twin.equals(gpu)
    assert twin.serialize() == blob  # deterministic encoding of equal state
    a, b = pair_table(gpu), pair_table(twin)
    assert len(a) > 20 and a.tobytes() == b.tobytes()

    movable = ids[body != 0]
    twin.move_translate(movable[:1], np.float32([[0.25, 0.0]]))
    assert not gpu.equals(twin)
    gpu.move_translate(movable[:1], np.float32([[0.25, 0.0]]))
    assert gpu.equals(twin)


@needs_ref
def test_grid_snapshots_travel_too():
    """PartitioningMethod::Grid: BroadPhaseTree images (per-cell DynamicBVHs, proxy -> cell handles) in both directions."""
    ref = H.Backend("reference", 1, 16)
    sc, ids, body, shape_ids = build_history(ref, seed=21)
    idx = P.SceneIndex(sc, shape_ids)
    gpu = H.Backend.deserialize("b200", ref.serialize(), method=1, entities=(ids, body, ref.get_prev_transforms(ids)))
    assert gpu.size() == ref.size()
    rng = np.random.default_rng(2)
    movable = ids[body != 0]
    for step in range(3):
        # compare_pairs asserts every manifold within 1e-5 relative; the bullets of build_history rotate, and rotating
        # sweeps interpolate sin/cos in double precision here (< 1 ulp from glibc's sinf/cosf): a few
        # manifolds may differ in the last bit
        res = P.compare_pairs("reference", ref, idx, ref.get_colliding_pairs(), gpu.get_colliding_pairs())
        assert res["expected"] > 50 and res["not_bit_identical"] <= 3, (step, res)
        d = rng.uniform(-1.0, 1.0, (len(movable), 2)).astype(np.float32)
        for reg in (ref, gpu):
            reg.move_translate(movable, d)
    ref2 = H.Backend.deserialize("reference", gpu.serialize(), method=1, entities=(ids, body, ref.get_prev_transforms(ids)))
    assert ref2.size() == ref.size()
    for step in range(3):
        assert same_pairs(pair_table(ref), pair_table(ref2), unique=False), step
        d = rng.uniform(-1.5, 1.5, (len(movable), 2)).astype(np.float32)  # proxies change cells
        for reg in (ref, ref2):
            reg.move_translate(movable, d)
    twin = H.Backend.deserialize("b200", gpu.serialize(), method=1)
    assert gpu.equals(twin)


@needs_ref
def test_snapshots_whose_shape_table_ends_in_freed_ids():
    """the snapshot's shape table includes freed ids; the device arrays only ever grew to the highest LIVE id
    re-added.  (a) every shape removed: the device capacity is still 0 when the boxes are restored; (b) 5000 slots with
    all live ids below 4096 (the first capacity step).  Both round trips work in both directions."""
    # (a)
    for name in ("reference", "b200"):
        reg = H.Backend(name)
        reg.create_entities([1, 2], [H.DYNAMIC, H.STATIC], np.zeros((2, 3), np.float32))
        g = np.zeros((2, 16), np.float32)
        g[:, 2] = 1.0
        s = reg.add_shapes([1, 2], [0, 0], g)
        reg.remove_shapes(s)
        blob = reg.serialize()
        for target in ("reference", "b200"):
            back = H.Backend.deserialize(target, blob)
            assert back.size() == 2 and len(back.get_colliding_pairs()) == 0
            # the freed ids are handed out again, LIFO
            assert list(back.add_shapes([1, 2], [0, 0], g)) == list(reg.add_shapes([1, 2], [0, 0], g))
            reg.remove_shapes(s)
    # (b)
    n = 5000
    ids = np.arange(n, dtype=np.uint32)
    xf3 = np.zeros((n, 3), np.float32)
    xf3[:, 0] = np.arange(n) * 0.8
    g = np.zeros((n, 16), np.float32)
    g[:, 2] = 0.5
    results = {}
    for name in ("reference", "b200"):
        reg = H.Backend(name)
        reg.create_entities(ids, np.full(n, H.DYNAMIC, np.uint8), xf3)
        reg.add_shapes(ids, np.zeros(n, np.uint8), g)
        reg.remove_entities(ids[4000:])
        blob = reg.serialize()
        for target in ("reference", "b200"):
            back = H.Backend.deserialize(target, blob)
            p = back.get_colliding_pairs()  # same-tree orientation is the backend's own: unordered keys
            results[(name, target)] = sorted(zip(np.minimum(p["shapeA"], p["shapeB"]).tolist(),
                                                 np.maximum(p["shapeA"], p["shapeB"]).tolist()))
            assert back.size() == 4000
    first = results[("reference", "reference")]
    assert len(first) == 3999 and all(v == first for v in results.values())
