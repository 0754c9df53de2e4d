"""GPU: the per-entity queries — Registry::getCollisions, areColliding, firstHitRayCast — against
the oracle on a scene with all body types, categories / masks, inactive shapes and multi-shape entities."""
import numpy as np
import pytest

from colli2de_b200 import scenes
from tests import harness as H, parity as P

pytestmark = pytest.mark.gpu

SUBJECT = "b200"
ORACLES = [n for n in ("oracle", "reference") if H.available(n)]


def _scene():
    sc = scenes.mixed_scene(6000, 600, 360.0, seed=41)
    sc["ent_body"][::20] = 2
    sc["shp_cat"][::5] = 2
    sc["shp_cat"][::7] = 3
    sc["shp_mask"][::3] = 1
    sc["shp_mask"][1::11] = 2
    return sc


def _prepare(name, sc):
    b = H.Backend(name)
    scenes.populate(b, sc)
    # a second shape on some entities (multi-shape entities)
    extra = np.arange(0, 6000, 17, dtype=np.uint32)
    g = np.zeros((len(extra), 16), np.float32)
    g[:, 0], g[:, 2] = 1.5, 0.8
    b.add_shapes(extra, np.zeros(len(extra), np.uint8), g)
    rng = scenes.LCG(23)
    mov = np.nonzero(sc["ent_body"] != 0)[0].astype(np.uint32)
    b.move_translate(mov, scenes.jitter(rng, len(mov), 0.6))
    bul = np.nonzero(sc["ent_body"] == 2)[0].astype(np.uint32)
    b.move_translate(bul, scenes.jitter(rng, len(bul), 9.0))
    off = np.arange(3, 6600, 29, dtype=np.uint32)
    b.set_active(off, np.zeros(len(off), np.uint8))
    return b


@pytest.mark.parametrize("oracle", ORACLES)
def test_get_collisions_are_colliding_first_hit(oracle):
    sc = _scene()
    ref, gpu = _prepare(oracle, sc), _prepare(SUBJECT, sc)
    ids = np.concatenate([np.arange(0, 6600, 13), np.nonzero(sc["ent_body"] == 2)[0][:80]]).astype(np.uint32)
    total = 0
    for e in ids:
        rp, gp = P.sort_pairs(ref.get_collisions(int(e))), P.sort_pairs(gpu.get_collisions(int(e)))
        assert P.pair_keys(rp).tobytes() == P.pair_keys(gp).tobytes(), f"entity {e}: different collision set"
        assert P.manifold_bits_equal(rp, gp).all(), f"entity {e}: manifolds differ"
        total += len(rp)
    assert total > 300
    # areColliding on a mix of near and far entity pairs, all body-type combinations
    rs = np.random.default_rng(1)
    a = rs.integers(0, 6600, 4000).astype(np.uint32)
    near = np.concatenate([[p["entityA"], p["entityB"]] for p in ref.get_colliding_pairs()[:1500]]).reshape(-1, 2)
    ia = np.concatenate([a, near[:, 0].astype(np.uint32), near[:, 1].astype(np.uint32)])
    ib = np.concatenate([rs.integers(0, 6600, 4000).astype(np.uint32), near[:, 1].astype(np.uint32), near[:, 0].astype(np.uint32)])
    ra, ga = ref.are_colliding(ia, ib), gpu.are_colliding(ia, ib)
    assert np.array_equal(ra, ga) and ra.sum() > 1000 and (ra == 0).sum() > 1000
    # firstHitRayCast
    ray4, inf = scenes.rays(2000, 360.0, seed=3)
    masks = np.full(2000, ~np.uint64(0))
    masks[::4] = 1
    rf, rh = ref.first_hit_raycast(ray4, inf, masks)
    gf, gh = gpu.first_hit_raycast(ray4, inf, masks)
    assert rf.sum() > 800
    bad = int((rf != gf).sum() + (rh[rf & gf > 0].tobytes() != gh[rf & gf > 0].tobytes()))
    if oracle == "oracle":
        assert bad == 0
    else:  # exact (entry, exit) ties between leaf boxes are resolved by the reference's traversal order (D10)
        differing = int((rf != gf).sum()) + int(sum(x.tobytes() != y.tobytes() for x, y in zip(rh[rf & gf > 0], gh[rf & gf > 0])))
        assert differing <= 10, differing


def test_get_collisions_batch_equals_the_single_calls(monkeypatch):
    """Registry::getCollisions(span of ids) (extension): one device pass, the concatenation of the per-entity results."""
    sc = _scene()
    gpu = _prepare(SUBJECT, sc)
    ids = np.arange(0, 6600, 7, dtype=np.uint32)
    singles = np.concatenate([gpu.get_collisions(int(e)) for e in ids])
    monkeypatch.setenv("RN_COLLISIONS_BATCH", "1")
    total, _ = gpu.get_collisions_many(ids)
    batch = np.zeros(total, H.PAIR_DTYPE)
    if total:
        gpu.lib.rn_copy_pairs(gpu.ctx, batch.ctypes.data_as(__import__("ctypes").c_void_p))
    assert total == len(singles) > 100
    assert P.sort_pairs(batch).tobytes() == P.sort_pairs(singles).tobytes()


@pytest.mark.parametrize("oracle", ORACLES)
def test_get_collisions_memoised_between_mutations(oracle):
    """After a few single getCollisions() calls without a mutation the Registry fetches the collisions of ALL entities in
    one device pass and serves further calls from that table (c2d_gpu_mutation_epoch); any mutation invalidates it."""
    sc = _scene()
    ref, gpu = _prepare(oracle, sc), _prepare(SUBJECT, sc)
    ids = np.arange(0, 6600, 11, dtype=np.uint32)

    def check(entities):
        for e in entities.tolist():
            a, b = P.sort_pairs(ref.get_collisions(e)), P.sort_pairs(gpu.get_collisions(e))
            assert np.array_equal(P.pair_keys(a), P.pair_keys(b)), e
            assert P.manifold_close(a, b).all(), e

    check(ids[:40])  # calls 5.. come out of the table
    total, sec = gpu.get_collisions_many(ids)  # 600 more calls, all table lookups now
    assert total > 100 and sec < 0.02, (total, sec)  # 600 device round trips would take ~18 ms
    # a move invalidates the table: the mover's neighbours see it at once
    mover = int(ids[3])
    target = ref.get_transforms(np.uint32([int(ids[5])]))[0, :2]
    for b in (ref, gpu):
        b.teleport_translation(np.uint32([mover]), target[None, :])
    check(np.uint32([mover, int(ids[5])]))
    check(ids[:12])
    # an entity created after the table was built, with and without shapes
    for b in (ref, gpu):
        b.create_entities(np.uint32([777001]), np.uint8([1]), np.float32([[5.0, 5.0, 0.0]]))
    check(ids[:8])
    assert len(gpu.get_collisions(777001)) == 0
    g = np.zeros((1, 16), np.float32)
    g[0, 2] = 30.0
    for b in (ref, gpu):
        b.add_shapes(np.uint32([777001]), np.uint8([0]), g)
    check(np.uint32([777001]))
    assert len(gpu.get_collisions(777001)) > 3
