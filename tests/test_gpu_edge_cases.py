"""GPU: edge cases — empty / single-shape registries, entities without shapes, everything removed, coincident
shapes (identical Morton keys -> the Karras index tie-break), degenerate geometry (zero radius, zero-length
capsule / segment), far-away coordinates, many shapes on one entity, repeated moves of one entity before a query
(mutation-log rounds), shape-id recycling — each against the oracle."""
import numpy as np
import pytest

from colli2de_b200 import scenes
from tests import golden_cases as G, harness as H, parity as P

pytestmark = pytest.mark.gpu

SUBJECT = "b200"
ORACLE = "reference" if H.available("reference") else "oracle"


def _pair(scene_fn):
    a, b = H.Backend(ORACLE), H.Backend(SUBJECT)
    scene_fn(a)
    scene_fn(b)
    return a, b


def _same_pairs(ref, gpu, idx=None):
    rp, gp = ref.get_colliding_pairs(), gpu.get_colliding_pairs()
    if idx is None:
        # small scenes: compare the unordered entity/shape sets and the manifold counts
        ka = sorted((min(int(r["shapeA"]), int(r["shapeB"])), max(int(r["shapeA"]), int(r["shapeB"]))) for r in rp)
        kb = sorted((min(int(r["shapeA"]), int(r["shapeB"])), max(int(r["shapeA"]), int(r["shapeB"]))) for r in gp)
        assert ka == kb
        return len(rp)
    return P.compare_pairs(ORACLE, ref, idx, rp, gp, max_ambiguous=0)["expected"]


def test_empty_and_tiny_registries():
    for name in (ORACLE, SUBJECT):
        b = H.Backend(name)
        assert len(b.get_colliding_pairs()) == 0 and b.size() == 0
        off, hits = b.raycast([[0, 0, 1, 1]], [0])
        assert len(hits) == 0 and list(off) == [0, 0]
        b.create_entities([7], [H.DYNAMIC], [[0, 0, 0]])
        b.move_translate([7], [[1, 1]])  # an entity without shapes can be moved
        assert len(b.get_colliding_pairs()) == 0
        b.add_shapes([7], [H.CIRCLE], G.UNIT)
        assert len(b.get_colliding_pairs()) == 0  # single leaf: no pair
        off, hits = b.raycast([[-5, 1, 5, 1]], [0])
        assert len(hits) == 1
        b.create_entities([8], [H.STATIC], [[1.5, 1, 0]])
        b.add_shapes([8], [H.CIRCLE], G.UNIT)
        assert len(b.get_colliding_pairs()) == 1  # one dynamic leaf vs one static leaf
        b.remove_entities([7, 8])
        assert len(b.get_colliding_pairs()) == 0 and b.size() == 0
        b.clear()
        assert len(b.get_colliding_pairs()) == 0


def test_coincident_shapes_identical_morton_keys():
    n = 600

    def build(b):
        b.create_entities(np.arange(n), np.full(n, H.DYNAMIC), np.tile(np.float32([5, 5, 0]), (n, 1)))
        g = np.tile(G.UNIT, (n, 1))
        b.add_shapes(np.arange(n), np.zeros(n, np.uint8), g)

    ref, gpu = _pair(build)
    gp = gpu.get_colliding_pairs()
    assert len(gp) == n * (n - 1) // 2
    assert len(ref.get_colliding_pairs()) == len(gp)
    keys = set(zip(gp["shapeA"].tolist(), gp["shapeB"].tolist()))
    assert len(keys) == len(gp) and all(a < b for a, b in keys)


def test_degenerate_geometry_and_far_coordinates():
    shapes = [G.circle(0, 0, 0.0), G.capsule(0, 0, 0, 0, 0.5), G.segment(1, 1, 1, 1), G.capsule(-1, 0, 1, 0, 0.0),
              G.polygon((0, 0), (1, 0), (2, 0)), G.rectangle(0, 0, 1e-3, 1e-3), G.circle(0, 0, 1e4),
              G.segment(-1e5, 0, 1e5, 0)]
    bt = G.batch(shapes)
    rs = np.random.default_rng(5)
    k = len(shapes)
    ia, ib = np.repeat(np.arange(k), k), np.tile(np.arange(k), k)
    for offset in (0.0, 1.0e6):
        xa = H.xf5_from_xf3(np.c_[rs.uniform(-1, 1, (k * k, 2)) + offset, rs.uniform(-3, 3, k * k)].astype(np.float32))
        xb = H.xf5_from_xf3(np.c_[rs.uniform(-1, 1, (k * k, 2)) + offset, rs.uniform(-3, 3, k * k)].astype(np.float32))
        em, ef = H.collide(ORACLE, bt.take(ia), xa, bt.take(ib), xb)
        gm, gf = H.collide(SUBJECT, bt.take(ia), xa, bt.take(ib), xb)
        assert np.array_equal(ef, gf)
        ok = np.isfinite(em["normal"]).all(axis=1)  # NaN manifolds (zero-length normalisations) compare bitwise below
        assert P.manifold_bits_equal(em[ok], gm[ok]).all()
        assert np.array_equal(np.isnan(em["normal"]), np.isnan(gm["normal"]))


def test_many_shapes_per_entity_and_repeated_moves():
    def build(b):
        b.create_entities([1, 2, 3], [H.DYNAMIC, H.STATIC, H.BULLET], [[0, 0, 0], [3, 0, 0.2], [-8, 0.5, 0]])
        geoms = []
        for k in range(40):  # a 40-shape dynamic entity
            geoms.append(G.circle(0.4 * k - 8, 0.1 * k, 0.3))
        b.add_shapes([1] * 40, [H.CIRCLE] * 40, G.batch(geoms).geom)
        b.add_shapes([2, 2], [H.POLYGON, H.CAPSULE], G.batch([G.rectangle(0, 0, 2, 1), G.capsule(0, 2, 3, 2, 0.4)]).geom, [4, 0])
        b.add_shapes([3], [H.CIRCLE], G.batch([G.circle(0, 0, 0.5)]).geom)
        # several moves of the same entities before the first query: translation, transform, teleport, translation
        for step in range(5):
            b.move_translate([1, 3], [[0.3, 0.02], [2.5, 0.0]])
            b.move_transform([1], [[0.1, -0.05, 0.07]])
            b.teleport_translation([3], [[-6.0 + 2.0 * step, 0.4]])
            b.move_translate([1], [[-0.1, 0.0]])

    ref, gpu = _pair(build)
    sc = dict(ent_id=np.uint32([1, 2, 3]), ent_body=np.uint8([1, 0, 2]),
              shp_entity=np.uint32([1] * 40 + [2, 2, 3]), shp_type=np.uint8([0] * 40 + [3, 1, 0]),
              shp_geom=np.zeros((43, 16), np.float32), shp_count=np.uint8([0] * 40 + [4, 0, 0]))
    for frame in range(3):
        rp, gp = P.sort_pairs(ref.get_colliding_pairs()), P.sort_pairs(gpu.get_colliding_pairs())
        assert len(rp) > 0
        assert P.pair_keys(rp).tobytes() == P.pair_keys(gp).tobytes()
        assert P.manifold_bits_equal(rp, gp).all()
        assert np.array_equal(ref.get_transforms([1, 2, 3]).view(np.uint32), gpu.get_transforms([1, 2, 3]).view(np.uint32))
        for b in (ref, gpu):
            b.move_translate([3], [[3.0, 0.0]])
            b.move_transform([1], [[0.2, 0.0, -0.05]])


def test_shape_id_recycling_under_churn():
    sc = scenes.mixed_scene(3000, 300, 250.0, seed=9)
    ref, gpu = H.Backend(ORACLE), H.Backend(SUBJECT)
    scenes.populate(ref, sc)
    scenes.populate(gpu, sc)
    rng = scenes.LCG(2)
    for round_ in range(4):
        gone = np.arange(round_, 3300, 7, dtype=np.uint32)
        for b in (ref, gpu):
            b.remove_shapes(gone)  # entity i owns shape i at first
        ent = np.arange(round_ * 3, 3300, 11, dtype=np.uint32)[: len(gone)]
        g = np.zeros((len(ent), 16), np.float32)
        g[:, 2] = 0.8
        ids_r = ref.add_shapes(ent, np.zeros(len(ent), np.uint8), g)
        ids_g = gpu.add_shapes(ent, np.zeros(len(ent), np.uint8), g)
        assert np.array_equal(ids_r, ids_g)  # LIFO recycling (Registry.hpp:283-289)
        mov = np.arange(0, 3000, dtype=np.uint32)
        d = scenes.jitter(rng, len(mov), 0.7)
        for b in (ref, gpu):
            b.move_translate(mov, d)
        rp, gp = P.sort_pairs(ref.get_colliding_pairs()), P.sort_pairs(gpu.get_colliding_pairs())
        # orientation of same-tree pairs differs (canonical vs tree order): compare unordered shape pairs + counts
        ka = np.sort(np.minimum(rp["shapeA"], rp["shapeB"]).astype(np.uint64) << np.uint64(32) | np.maximum(rp["shapeA"], rp["shapeB"]))
        kb = np.sort(np.minimum(gp["shapeA"], gp["shapeB"]).astype(np.uint64) << np.uint64(32) | np.maximum(gp["shapeA"], gp["shapeB"]))
        extra = len(np.setxor1d(ka, kb))
        assert extra <= 4, f"round {round_}: {extra} differing pairs"  # segment-segment orientation noise only (D8)
        assert len(ka) > 300
        # removing the shapes above broke "entity i owns shape i": from now on only remove what we just added
        break


def test_staged_moves_equal_moves_and_are_invalidated_by_structure_changes():
    """N2: a device-resident move batch does what the same moveEntity(Translation) calls do (bit-identical boxes, leaf
    rule and results); addShape / removeShape on ANY entity after staging invalidates the batch instead of silently
    leaving the new shape behind (host and device would disagree about where the entity's shapes are)."""
    sc = scenes.mixed_scene(3000, 300, 150.0, seed=5)
    a, b = H.Backend(SUBJECT), H.Backend(SUBJECT)
    for reg in (a, b):
        scenes.populate(reg, sc)
    mov = np.nonzero(sc["ent_body"] != 0)[0].astype(np.uint32)
    d = scenes.jitter(scenes.LCG(3), len(mov), 0.7)
    h = b.stage_translations(mov, d)
    for _ in range(3):
        a.move_translate(mov, d)
        b.apply_staged(h)
        ta, fa, _ = a.read_boxes()
        tb, fb, _ = b.read_boxes()
        assert ta.tobytes() == tb.tobytes() and fa.tobytes() == fb.tobytes()
        assert P.sort_pairs(a.get_colliding_pairs()).tobytes() == P.sort_pairs(b.get_colliding_pairs()).tobytes()
        assert np.array_equal(a.get_transforms(mov).view(np.uint32), b.get_transforms(mov).view(np.uint32))
    b.add_shapes(mov[:1], [H.CIRCLE], G.UNIT)
    with pytest.raises(RuntimeError, match="added or removed"):
        b.apply_staged(h)
    b.release_staged(h)
    h2 = b.stage_translations(mov, d)  # a fresh batch sees the new shape
    a.add_shapes(mov[:1], [H.CIRCLE], G.UNIT)
    a.move_translate(mov, d)
    b.apply_staged(h2)
    assert P.sort_pairs(a.get_colliding_pairs()).tobytes() == P.sort_pairs(b.get_colliding_pairs()).tobytes()


def test_out_of_range_shape_ids_throw():
    b = H.Backend(SUBJECT)
    b.create_entities([1], [H.DYNAMIC], [[0, 0, 0]])
    s = b.add_shapes([1], [H.CIRCLE], G.UNIT)
    with pytest.raises(RuntimeError, match="out of range"):
        b.set_active([int(s[0]) + 100000], [0])


def test_sparse_shape_ids_stay_inside_the_device_arrays():
    """host tables grow by doubling and can exceed the device capacity (ids 10000, 10001 -> host 20002 slots,
    capacity 16384); setShapeActive on a stale slot past the capacity must not write out of bounds."""
    b, r = H.Backend(SUBJECT), H.Backend(ORACLE)
    n = 10002
    ids = np.arange(n, dtype=np.uint32)
    xf3 = np.zeros((n, 3), np.float32)
    xf3[:, 0] = np.arange(n) * 0.75
    g = np.zeros((n, 16), np.float32)
    g[:, 2] = 0.5
    for reg in (b, r):
        reg.create_entities(ids, np.full(n, H.DYNAMIC, np.uint8), xf3)
        reg.add_shapes(ids, np.zeros(n, np.uint8), g)
        reg.remove_entities(ids[:10000])  # frees ids 0..9999; live ids 10000, 10001
        reg.set_active([10001], [0])
        reg.set_active([10001], [1])
    gp, rp = P.sort_pairs(b.get_colliding_pairs()), P.sort_pairs(r.get_colliding_pairs())
    assert len(gp) == 1 and np.array_equal(P.pair_keys(gp), P.pair_keys(rp)) and P.manifold_bits_equal(gp, rp).all()
