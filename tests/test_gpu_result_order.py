"""GPU: c2d_gpu_set_result_order / Registry::setDeterministicOrder - the records come back in the reference's order
(Registry.hpp:487-545: five body-type groups one after the other, each sorted by (shapeA, shapeB)) and are byte-identical
from run to run; cross-tree groups equal the reference's own sequence record for record."""
import numpy as np
import pytest

from colli2de_b200 import capi, scenes
from tests import harness as H, parity as P

pytestmark = pytest.mark.gpu

ORACLE = "reference" if H.available("reference") else "oracle"


def _group(idx, p):
    ba, bb = idx.body[p["shapeA"]].astype(np.int64), idx.body[p["shapeB"]].astype(np.int64)
    return np.where(ba == 1, np.where(bb == 0, 0, 1), np.where(bb == 0, 2, np.where(bb == 1, 3, 4)))


@pytest.mark.parametrize("n_dyn,small", [(30000, False), (3000, True)])
def test_reference_order(n_dyn, small):
    sc = scenes.mixed_scene(n_dyn, n_dyn // 10, 900.0 * (n_dyn / 30000.0) ** 0.5, seed=21)
    sc["ent_body"][::30] = 2
    ref, a, b = H.Backend(ORACLE), H.Backend("b200"), H.Backend("b200")
    for x in (ref, a, b):
        scenes.populate(x, sc)
    for x in (a, b):
        capi.set_result_order(x.native_context(), True)
    idx = P.SceneIndex(sc)
    rng = scenes.LCG(8)
    mov = np.nonzero(sc["ent_body"] != 0)[0].astype(np.uint32)
    for frame in range(3):
        d = scenes.jitter(rng, len(mov), 0.7)
        for x in (ref, a, b):
            x.move_translate(mov, d)
        ra, rb, rr = a.get_colliding_pairs(), b.get_colliding_pairs(), ref.get_colliding_pairs()
        assert len(ra) > 500
        # run to run: two registries with the same history return byte-identical arrays (records compared by fields: the
        # 84-byte record has 3 padding bytes)
        for f in ra.dtype.names:
            assert np.array_equal(ra[f], rb[f]) if ra[f].dtype.names is None else ra[f].tobytes() == rb[f].tobytes(), f
        # globally ordered by (group, shapeA, shapeB)
        g = _group(idx, ra)
        key = (g.astype(np.uint64) << np.uint64(58)) | (ra["shapeA"].astype(np.uint64) << np.uint64(29)) | ra["shapeB"].astype(np.uint64)
        assert (np.diff(key.astype(np.int64)) > 0).all(), "records are not in (group, shapeA, shapeB) order"
        # cross-tree groups: the reference's own sequence, record for record
        gr = _group(idx, rr)
        for grp in (0, 2, 3):
            mine, theirs = ra[g == grp], rr[gr == grp]
            assert len(mine) == len(theirs)
            for f in ("entityA", "entityB", "shapeA", "shapeB"):
                assert np.array_equal(mine[f], theirs[f]), (grp, f)
        # and the set as a whole is still the reference's
        P.compare_pairs(ORACLE, ref, idx, rr, ra, max_ambiguous=8)
    # switching the order off again returns the plain (unordered) result of the same frame
    capi.set_result_order(a.native_context(), False)
    plain = a.get_colliding_pairs()
    assert P.sort_pairs(plain)["shapeA"].tobytes() == P.sort_pairs(ra)["shapeA"].tobytes()
