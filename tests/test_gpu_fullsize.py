"""GPU: BASELINE.json's full sizes.  Where the reference still finishes in seconds on the host the comparison is
direct (exact sorted tuple set + manifolds); on top of that, size-independent properties: no duplicate pair,
canonical orientation, idempotence of the query, agreement of two independent broadphases (LBVH vs uniform grid),
and spot re-evaluation of reported / rejected candidates through the stateless narrowphase."""
import time

import numpy as np
import pytest

from colli2de_b200 import capi, scenes
from tests import harness as H, parity as P

pytestmark = pytest.mark.gpu

SUBJECT = "b200"
HAVE_REF = H.available("reference")
_T0 = [time.time()]


def _lap(label):
    now = time.time()
    print(f"[{label}: {now - _T0[0]:.1f}s]", flush=True)
    _T0[0] = now


def _keys(p):
    return (p["shapeA"].astype(np.uint64) << np.uint64(32)) | p["shapeB"].astype(np.uint64)


def test_cfg3_one_million_mixed_shapes():
    sc = scenes.mixed_scene(1_000_000, 100_000, 5000.0, clustered=True)
    n_dyn = 1_000_000
    idx = P.SceneIndex(sc)
    gpu, grid = H.Backend(SUBJECT), H.Backend(SUBJECT, 1, 64)
    _lap("cfg3 scene")
    scenes.populate(gpu, sc)
    scenes.populate(grid, sc)
    _lap("cfg3 populate gpu x2")
    rng = scenes.LCG(2024)
    dyn = np.arange(n_dyn, dtype=np.uint32)
    d = scenes.jitter(rng, n_dyn, 0.5)
    gpu.move_translate(dyn, d)
    grid.move_translate(dyn, d)
    p1 = gpu.get_colliding_pairs()
    st = gpu.stats()
    assert st["retries"] <= 2 and st["candidates"] >= len(p1) > 300_000
    # properties
    k = _keys(p1)
    assert len(np.unique(k)) == len(k), "duplicate pairs"
    same = idx.body[p1["shapeA"]] == idx.body[p1["shapeB"]]
    assert (p1["shapeA"][same] < p1["shapeB"][same]).all()
    assert (idx.body[p1["shapeA"][~same]] == H.DYNAMIC).all() and (idx.body[p1["shapeB"][~same]] == H.STATIC).all()
    assert (p1["entityA"] != p1["entityB"]).all() and (p1["pointCount"] >= 1).all() and (p1["pointCount"] <= 2).all()
    p2 = gpu.get_colliding_pairs()  # idempotent: nothing moved
    assert P.sort_pairs(p1).tobytes() == P.sort_pairs(p2).tobytes()
    _lap("cfg3 gpu queries")
    pg = grid.get_colliding_pairs()  # independent broadphase (uniform grid), same narrowphase
    _lap("cfg3 grid query")
    assert P.sort_pairs(p1).tobytes() == P.sort_pairs(pg).tobytes()
    # spot check: re-evaluate a sample of reported pairs and of rejected candidates through collide_batch
    cand = capi.read_candidates(gpu.native_context())
    rs = np.random.default_rng(0)
    xf = gpu.get_transforms(np.arange(len(sc["ent_id"]), dtype=np.uint32))
    sel = rs.choice(len(p1), 20000, replace=False)
    a, b = p1["shapeA"][sel].astype(np.int64), p1["shapeB"][sel].astype(np.int64)
    man, flag = H.collide(SUBJECT, idx.batch(a), xf[a], idx.batch(b), xf[b])
    assert flag.all() and P.manifold_bits_equal(man, p1[sel]).all()
    ck = (cand[:, 0].astype(np.uint64) << np.uint64(32)) | cand[:, 1].astype(np.uint64)
    rejected = cand[~np.isin(ck, k)]
    assert len(rejected) > 100_000
    sel = rs.choice(len(rejected), 20000, replace=False)
    a, b = rejected[sel, 0].astype(np.int64), rejected[sel, 1].astype(np.int64)
    _, flag = H.collide(SUBJECT, idx.batch(a), xf[a], idx.batch(b), xf[b])
    assert not flag.any()
    _lap("cfg3 spot checks")
    if HAVE_REF:  # the reference needs ~1.5 s for this frame: compare directly
        ref = H.Backend("reference")
        scenes.populate(ref, sc)
        ref.move_translate(dyn, d)
        _lap("cfg3 reference populate+move")
        # observed: 5 orientation-ambiguous segment pairs in 415 k; the bound is ~2x that, so a real
        # regression of a few dozen same-tree pairs cannot hide behind the adjudication
        res = P.compare_pairs("reference", ref, idx, ref.get_colliding_pairs(), p1, max_ambiguous=12)
        assert res["not_bit_identical"] == 0, res
        _lap("cfg3 reference query + compare")
        print("cfg3 full size vs reference:", res)
        # the bench runs 25 frames of this scene, not one: four more frames of jitter, every frame compared, and the
        # secondary oracles (tight boxes, leaf boxes, candidate set) bit for bit at the end
        from tests.test_gpu_secondary_oracles import check_state

        for frame in range(1, 5):
            d = scenes.jitter(rng, n_dyn, 0.5)
            gpu.move_translate(dyn, d)
            ref.move_translate(dyn, d)
            res = P.compare_pairs("reference", ref, idx, ref.get_colliding_pairs(), gpu.get_colliding_pairs(),
                                  max_ambiguous=12)
            assert res["not_bit_identical"] == 0 and res["expected"] > 300_000, res
            print("cfg3 frame", frame, res)
        _lap("cfg3 four more frames vs reference")
        live, cands = check_state(ref, gpu, "cfg3 frame 4")
        assert live == 1_100_000 and cands > 1_000_000
        _lap("cfg3 boxes + candidates vs reference")
        print("cfg3 secondary oracles: shapes", live, "candidates", cands)


def test_cfg2_hundred_thousand_mixed_shapes():
    """BASELINE cfg 2 at its stated size: 100 k mixed dynamic + 10 k static polygons, uniform density, five frames."""
    if not HAVE_REF:
        pytest.skip("oracle/_ref (the compiled reference) is not available")
    sc = scenes.mixed_scene(100_000, 10_000, 1600.0)
    idx = P.SceneIndex(sc)
    gpu, ref = H.Backend(SUBJECT), H.Backend("reference")
    scenes.populate(gpu, sc)
    scenes.populate(ref, sc)
    rng = scenes.LCG(4242)
    dyn = np.arange(100_000, dtype=np.uint32)
    total = 0
    for frame in range(5):
        d = scenes.jitter(rng, len(dyn), 0.5)
        gpu.move_translate(dyn, d)
        ref.move_translate(dyn, d)
        res = P.compare_pairs("reference", ref, idx, ref.get_colliding_pairs(), gpu.get_colliding_pairs(), max_ambiguous=4)
        assert res["not_bit_identical"] == 0 and res["expected"] > 25_000, res
        total += res["expected"]
    print("cfg2 full size vs reference: pairs over 5 frames", total)
    _lap("cfg2 five frames")


def test_cfg4_bullets_against_one_million_static_polygons():
    _lap("start cfg4")
    sc = scenes.bullet_scene(1_000_000, 10_000, 3200.0)
    idx = P.SceneIndex(sc)
    gpu = H.Backend(SUBJECT)
    scenes.populate(gpu, sc)
    _lap("cfg4 scene + populate gpu")
    bul = np.arange(1_000_000, 1_010_000, dtype=np.uint32)
    rng = scenes.LCG(99)
    backends = [gpu]
    if HAVE_REF:
        ref = H.Backend("reference")
        scenes.populate(ref, sc)
        backends.append(ref)
        _lap("cfg4 reference populate")
    for frame in range(2):
        d = scenes.jitter(rng, len(bul), 20.0)
        for b in backends:
            b.move_translate(bul, d)
        p = gpu.get_colliding_pairs()
        assert len(p) > 10_000
        k = _keys(p)
        assert len(np.unique(k)) == len(k)
        assert (idx.body[p["shapeA"]] == H.BULLET).all()
        if HAVE_REF:
            res = P.compare_pairs("reference", ref, idx, ref.get_colliding_pairs(), p, max_ambiguous=50)
            assert res["not_bit_identical"] == 0, res
            _lap("cfg4 frame")
            print("cfg4 frame", frame, res)


def test_cfg5_rays_against_four_million_shapes():
    _lap("start cfg5")
    sc = scenes.mixed_scene(3_640_000, 360_000, 9535.0, seed=77)
    gpu = H.Backend(SUBJECT)
    scenes.populate(gpu, sc)
    _lap("cfg5 scene + populate gpu")
    n = 100_000
    ray4, inf = scenes.rays(n, 9535.0, seed=13)
    off, hits = gpu.raycast(ray4, inf)
    _lap("cfg5 gpu raycast")
    assert len(hits) > n
    # per-ray ordering and key uniqueness on a sample
    for i in range(0, n, 4001):
        h = hits[int(off[i]):int(off[i + 1])]
        keys = list(zip(h["entryTime"].tolist(), h["exitTime"].tolist()))
        assert keys == sorted(set(keys))
    # linearity of the hit points: entry = start + d * entryTime
    ri = np.repeat(np.arange(n), np.diff(off.astype(np.int64)))
    d = np.where(inf[ri, None] == 1, ray4[ri, 2:], ray4[ri, 2:] - ray4[ri, :2])
    assert np.array_equal(hits["entry"], ray4[ri, :2] + d * hits["entryTime"][:, None])
    assert (hits["entryTime"] <= hits["exitTime"]).all() and (hits["entryTime"] >= 0).all()
    assert (hits["exitTime"][inf[ri] == 0] <= 1.0).all()
    if HAVE_REF:  # the reference casts ~1 ms per ray at this size: compare 1 500 rays directly
        ref = H.Backend("reference")
        _lap("cfg5 property checks")
        scenes.populate(ref, sc)
        _lap("cfg5 reference populate")
        m = 1500
        eo, eh = ref.raycast(ray4[:m], inf[:m])
        _lap("cfg5 reference raycast")
        bad = 0
        for i in range(m):
            a, b = eh[int(eo[i]):int(eo[i + 1])], hits[int(off[i]):int(off[i + 1])]
            bad += not (len(a) == len(b) and a.tobytes() == b.tobytes())
        assert bad <= 3, f"{bad} of {m} rays differ"
        print("cfg5 rays vs reference:", bad, "of", m, "differ; hits", len(eh))
