"""GPU parity of Registry::getCollidingPairs through the drop-in c2d::Registry (runner compiled against
include/colli2de) -> C ABI -> CUDA, against the reference replaying the same scripted scene.  The pair set
must be exact as sorted (entityA, shapeA, entityB, shapeB) tuples (canonical orientation) and
manifolds within 1e-5 relative."""
import numpy as np
import pytest

from colli2de_b200 import scenes
from tests import harness as H, parity as P

pytestmark = pytest.mark.gpu

ORACLE = "reference" if H.available("reference") else "oracle"
SUBJECT = "b200"


def _both(scene, method=0, cell=120):
    ref, gpu = H.Backend(ORACLE, method, cell), H.Backend(SUBJECT, method, cell)
    ids_r = scenes.populate(ref, scene)
    ids_g = scenes.populate(gpu, scene)
    assert np.array_equal(ids_r, ids_g)
    return ref, gpu, P.SceneIndex(scene, ids_r)


def _check(ref, gpu, idx, max_ambiguous=None):
    rp = ref.get_colliding_pairs()
    gp = gpu.get_colliding_pairs()
    return P.compare_pairs(ORACLE, ref, idx, rp, gp, max_ambiguous)


def test_baseline_cfg1_static_frames():
    sc = scenes.baseline_cfg1()
    ref, gpu, idx = _both(sc)
    rng = scenes.LCG(5)
    dyn = np.nonzero(sc["ent_body"] == 1)[0].astype(np.uint32)
    for frame in range(4):
        res = _check(ref, gpu, idx, max_ambiguous=0)
        assert res["expected"] > 10000 and res["not_bit_identical"] == 0, res
        d = scenes.jitter(rng, len(dyn), 0.5)
        ref.move_translate(dyn, d)
        gpu.move_translate(dyn, d)


def test_mixed_with_bullets_translate_frames():
    sc = scenes.mixed_scene(30000, 3000, 900.0)
    sc["ent_body"][::40] = 2  # some bullets of every shape type
    ref, gpu, idx = _both(sc)
    rng = scenes.LCG(6)
    mov = np.nonzero(sc["ent_body"] != 0)[0].astype(np.uint32)
    bul = np.nonzero(sc["ent_body"] == 2)[0].astype(np.uint32)
    for frame in range(5):
        d = scenes.jitter(rng, len(mov), 0.6)
        ref.move_translate(mov, d)
        gpu.move_translate(mov, d)
        big = scenes.jitter(rng, len(bul), 15.0)
        ref.move_translate(bul, big)
        gpu.move_translate(bul, big)
        res = _check(ref, gpu, idx, max_ambiguous=8)
        assert res["expected"] > 5000 and res["not_bit_identical"] == 0, res
    assert np.array_equal(ref.get_transforms(mov).view(np.uint32), gpu.get_transforms(mov).view(np.uint32))


def test_transform_moves_teleports_removal_activation():
    sc = scenes.mixed_scene(8000, 800, 450.0, seed=99)
    sc["ent_body"][::25] = 2
    ref, gpu, idx = _both(sc)
    rng = scenes.LCG(8)
    n = len(sc["ent_id"])
    ids = np.arange(n, dtype=np.uint32)
    nonbullet = ids[sc["ent_body"] != 2]
    for frame in range(4):
        # rotate + translate every non-bullet (Rotation::operator+= drift, D5), translate bullets
        d3 = np.zeros((len(nonbullet), 3), np.float32)
        d3[:, :2] = scenes.jitter(rng, len(nonbullet), 0.4)
        d3[:, 2] = rng.next_float(-0.3, 0.3, len(nonbullet))
        for b in (ref, gpu):
            b.move_transform(nonbullet, d3)
        bul = ids[sc["ent_body"] == 2]
        db = scenes.jitter(rng, len(bul), 8.0)
        for b in (ref, gpu):
            b.move_translate(bul, db)
        # teleport a slice (fresh sin/cos), twice in a row for a few (same shape twice before a flush)
        tel = nonbullet[frame::97]
        t3 = np.zeros((len(tel), 3), np.float32)
        t3[:, :2] = rng.next_vec2(0.0, 450.0, len(tel))
        t3[:, 2] = rng.next_float(0.0, 6.28, len(tel))
        for b in (ref, gpu):
            b.teleport_transform(tel, t3)
            b.teleport_translation(tel[:5], t3[:5, :2] + np.float32(1.0))
        # deactivate / reactivate and remove some shapes
        off = ids[(frame * 3) % 11::11][:200]
        for b in (ref, gpu):
            b.set_active(off, np.zeros(len(off), np.uint8))
        res = _check(ref, gpu, idx, max_ambiguous=8)
        assert res["expected"] > 1000 and res["not_bit_identical"] == 0, res
        for b in (ref, gpu):
            b.set_active(off, np.ones(len(off), np.uint8))
    assert np.array_equal(ref.get_transforms(ids).view(np.uint32), gpu.get_transforms(ids).view(np.uint32))
    # rotating bullets: the sweep's interpolated sin/cos are evaluated in double precision on the device (glibc's
    # sinf on the host) -> pair set still exact here, manifolds within 1e-5 but not necessarily bit-identical
    bul = ids[sc["ent_body"] == 2]
    t3 = np.zeros((len(bul), 3), np.float32)
    t3[:, :2] = ref.get_transforms(bul)[:, :2] + scenes.jitter(rng, len(bul), 6.0)
    t3[:, 2] = rng.next_float(0.0, 6.28, len(bul))
    for b in (ref, gpu):
        b.teleport_transform(bul, t3)
    res = _check(ref, gpu, idx, max_ambiguous=8)
    assert res["expected"] > 1000 and res["not_bit_identical"] <= 0.02 * res["expected"], res
    # removal: entities and shapes, then shape-id recycling
    gone = ids[5::13]
    for b in (ref, gpu):
        b.remove_entities(gone)
    keep = np.ones(n, bool)
    keep[gone] = False
    rp, gp = ref.get_colliding_pairs(), gpu.get_colliding_pairs()
    assert not np.isin(gp["entityA"], gone).any() and not np.isin(gp["entityB"], gone).any()
    res = P.compare_pairs(ORACLE, ref, idx, rp, gp, 8)
    assert res["expected"] > 500
    assert ref.size() == gpu.size()


def test_category_filter_and_same_entity():
    """(catA & catB) != 0 decides, the mask is ignored on this path; shapes of one entity
    never collide with each other (Registry.hpp:619-621)."""
    for name in (ORACLE, SUBJECT):
        b = H.Backend(name)
        b.create_entities([1, 2, 3], [H.DYNAMIC, H.DYNAMIC, H.STATIC], np.zeros((3, 3), np.float32))
        g = np.zeros((4, 16), np.float32)
        g[:, 2] = 1.0
        s = b.add_shapes([1, 1, 2, 3], [0, 0, 0, 0], g, cat=[1, 1, 2, 3], mask=[0, 0, 0, 0])
        p = P.sort_pairs(b.get_colliding_pairs())
        got = [(int(r["entityA"]), int(r["shapeA"]), int(r["entityB"]), int(r["shapeB"])) for r in p]
        # 1&2 = 0 -> no pair between entities 1 and 2; 3 matches both
        assert got == [(1, 0, 3, 3), (1, 1, 3, 3), (2, 2, 3, 3)], (name, got)


def test_bulk_move_and_teleport_entities():
    """Registry::moveEntities / teleportEntities (N2: one call per batch instead of one per entity) against the reference's
    single calls (Registry.hpp:403-411, 434-453): same pairs, same manifolds, same transforms, bullets' sweeps included;
    a batch that names an entity twice in a row applies both moves in order."""
    sc = scenes.mixed_scene(30000, 3000, 900.0, seed=9)
    sc["ent_body"][::40] = 2
    ref, gpu, idx = _both(sc)
    rng = scenes.LCG(16)
    mov = np.nonzero(sc["ent_body"] != 0)[0].astype(np.uint32)
    for frame in range(5):
        if frame % 2 == 0:
            d = scenes.jitter(rng, len(mov), 0.7)
            ref.move_translate(mov, d)  # moveEntity(id, Translation) per entity
            gpu.move_entities_bulk(mov, d)
        else:
            target = ref.get_transforms(mov)[:, :2] + scenes.jitter(rng, len(mov), 4.0)
            ref.teleport_translation(mov, target)
            gpu.move_entities_bulk(mov, target, teleport=True)
        if frame == 2:  # the same entities twice in one batch
            twice = np.concatenate([mov[:500], mov[:500]])
            d2 = scenes.jitter(rng, len(twice), 0.3)
            ref.move_translate(twice, d2)
            gpu.move_entities_bulk(twice, d2)
        res = _check(ref, gpu, idx, max_ambiguous=8)
        assert res["expected"] > 5000 and res["not_bit_identical"] == 0, res
        assert np.array_equal(ref.get_transforms(mov).view(np.uint32), gpu.get_transforms(mov).view(np.uint32))


def test_incremental_refit_many_frames_without_rebuild():
    """k_refit_dirty: with the topology kept for many frames, internal boxes only grow (conservative) while the leaf slots
    stay exact - the pair set must stay the reference's through large moves, Transform moves, teleports and bullets, and
    be identical to the full bottom-up fit's."""
    from colli2de_b200 import capi

    sc = scenes.mixed_scene(30000, 3000, 900.0, seed=31, clustered=True)
    sc["ent_body"][::50] = 2
    ref, gpu, idx = _both(sc)
    full = H.Backend(SUBJECT)
    scenes.populate(full, sc)
    capi.set_small_scene_limit(gpu.native_context(), 0)
    capi.set_rebuild_period(gpu.native_context(), 64)
    capi.set_rebuild_period(full.native_context(), 64)
    capi.set_incremental_refit(full.native_context(), False)
    rng = scenes.LCG(77)
    mov = np.nonzero(sc["ent_body"] != 0)[0].astype(np.uint32)
    for frame in range(12):
        d = scenes.jitter(rng, len(mov), 3.0 if frame % 3 else 0.4)
        for x in (ref, gpu, full):
            x.move_translate(mov, d)
        if frame in (4, 9):  # Transform moves and teleports flag their leaves as well
            some = np.nonzero(sc["ent_body"] == 1)[0].astype(np.uint32)[frame::17]  # (rotating bullets are tolerance-only)
            dxf = np.concatenate([scenes.jitter(rng, len(some), 2.0), scenes.jitter(rng, len(some), 0.3)[:, :1]], axis=1)
            for x in (ref, gpu, full):
                x.move_transform(some, dxf)
        res = _check(ref, gpu, idx, max_ambiguous=8)
        assert res["expected"] > 5000 and res["not_bit_identical"] == 0, (frame, res)
        a, b = P.sort_pairs(gpu.get_colliding_pairs()), P.sort_pairs(full.get_colliding_pairs())
        assert P.pair_keys(a).tobytes() == P.pair_keys(b).tobytes()
        assert np.array_equal(gpu.candidates(), full.candidates()) or \
            np.array_equal(np.sort(gpu.candidates().view("u8").ravel()), np.sort(full.candidates().view("u8").ravel()))
