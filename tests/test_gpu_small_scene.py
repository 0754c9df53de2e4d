"""The small-scene shortcut (k_small_pairs + k_small_narrow, c2d_gpu_set_small_scene_limit) must give exactly the
results of the tree pipeline and of the reference: same candidate predicate (leaf boxes intersect, categories match),
same pair orientation, same narrowphase."""
import numpy as np
import pytest

from colli2de_b200 import capi, scenes
from tests import harness as H, parity as P

pytestmark = pytest.mark.gpu

ORACLE = "reference" if H.available("reference") else "oracle"


def _sorted_records(recs):
    order = np.lexsort((recs["shapeB"], recs["shapeA"], recs["entityB"], recs["entityA"]))
    out = recs[order].copy()
    out["_pad"] = 0
    return out


@pytest.mark.parametrize("method,cell", [(0, 0), (1, 64)])
def test_small_path_equals_tree_path_and_reference(method, cell):
    sc = scenes.mixed_scene(2400, 400, 230.0, seed=31)
    sc["ent_body"][::13] = 2
    sc["shp_cat"][::7] = 2  # category filtering in the broadphase
    ref = H.Backend(ORACLE, method, cell or 120)
    small = H.Backend("b200", method, cell or 120)
    tree = H.Backend("b200", method, cell or 120)
    capi.set_small_scene_limit(small.native_context(), 1 << 20)
    capi.set_small_scene_limit(tree.native_context(), 0)
    ids = scenes.populate(ref, sc)
    for b in (small, tree):
        assert np.array_equal(scenes.populate(b, sc), ids)
    idx = P.SceneIndex(sc, ids)
    rng = scenes.LCG(3)
    mov = np.nonzero(sc["ent_body"] != 0)[0].astype(np.uint32)
    for frame in range(4):
        d = scenes.jitter(rng, len(mov), 0.8)
        for b in (ref, small, tree):
            b.move_translate(mov, d)
        a, t = small.get_colliding_pairs(), tree.get_colliding_pairs()
        assert small.stats()["kernel_launches"] <= 8 < tree.stats()["kernel_launches"]
        assert len(a) > 300 and _sorted_records(a).tobytes() == _sorted_records(t).tobytes(), frame
        res = P.compare_pairs(ORACLE, ref, idx, ref.get_colliding_pairs(), a)
        assert res["not_bit_identical"] == 0 and res["expected"] > 300, (frame, res)
    # rays still find their trees (built on demand, the small path keeps none)
    ray4, inf = scenes.rays(32, 230.0, seed=5, finite_len=40.0)
    if method == 0:
        o1, h1 = small.raycast(ray4, inf)
        o2, h2 = ref.raycast(ray4, inf)
        assert np.array_equal(o1, o2) and h1.tobytes() == h2.tobytes()


def test_scene_growing_past_the_limit_switches_pipelines():
    b = H.Backend("b200")
    capi.set_small_scene_limit(b.native_context(), 1500)
    ref = H.Backend(ORACLE)
    sc = scenes.mixed_scene(1200, 200, 160.0, seed=8)
    ids = scenes.populate(ref, sc)
    scenes.populate(b, sc)
    idx = P.SceneIndex(sc, ids)
    res = P.compare_pairs(ORACLE, ref, idx, ref.get_colliding_pairs(), b.get_colliding_pairs())
    assert res["not_bit_identical"] == 0 and b.stats()["kernel_launches"] <= 8
    # 600 more entities: past the limit, the tree pipeline takes over
    n0 = len(sc["ent_id"])
    extra = np.arange(n0, n0 + 600, dtype=np.uint32)
    xf3 = np.zeros((600, 3), np.float32)
    xf3[:, :2] = scenes.LCG(9).next_vec2(0.0, 160.0, 600)
    g = np.zeros((600, 16), np.float32)
    g[:, 2] = 1.0
    for reg in (ref, b):
        reg.create_entities(extra, np.ones(600, np.uint8), xf3)
        reg.add_shapes(extra, np.zeros(600, np.uint8), g)
    rp, gp = ref.get_colliding_pairs(), b.get_colliding_pairs()
    assert b.stats()["kernel_launches"] > 8
    key = lambda r: set(zip(np.minimum(r["shapeA"], r["shapeB"]).tolist(), np.maximum(r["shapeA"], r["shapeB"]).tolist()))
    assert key(rp) == key(gp) and len(gp) > 0
