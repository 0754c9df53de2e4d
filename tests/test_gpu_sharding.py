"""GPU: the multi-GPU decompositions on ONE device (ranks emulated as separate contexts, run one after the other):
  * strong: every rank mirrors the scene and owns a Morton-rank slab of the query leaves (c2d_gpu_set_shard);
  * weak: every rank holds only its spatial slab plus halo copies (c2d_gpu_shape_set_ghost + the owner rule).
In both the union of the ranks' pair lists must equal the single-registry result, each pair exactly once."""
import numpy as np
import pytest

from colli2de_b200 import capi, scenes, sharding
from tests import harness as H, parity as P

pytestmark = pytest.mark.gpu

SUBJECT = "b200"


def _entity_keys(p):
    a = np.minimum(p["entityA"], p["entityB"]).astype(np.uint64)
    b = np.maximum(p["entityA"], p["entityB"]).astype(np.uint64)
    return np.sort(a << np.uint64(32) | b)


def _scene():
    sc = scenes.mixed_scene(30000, 3000, 800.0, seed=19, clustered=True)
    sc["ent_body"][::35] = 2
    return sc


def _moves(sc, frames=3):
    rng = scenes.LCG(5)
    return [scenes.jitter(rng, len(sc["ent_id"]), 0.6) for _ in range(frames)]


@pytest.mark.parametrize("method,world", [(0, 2), (0, 5), (1, 3)])
def test_strong_query_slabs(method, world):
    sc = _scene()
    movable = sc["ent_body"] != 0
    full = H.Backend(SUBJECT, method, 24)
    scenes.populate(full, sc)
    ranks = []
    for r in range(world):
        b = H.Backend(SUBJECT, method, 24)
        scenes.populate(b, sc)
        capi.set_shard(b.native_context(), r, world)
        ranks.append(b)
    for d in _moves(sc):
        for b in [full] + ranks:
            b.move_translate(sc["ent_id"][movable], d[movable])
        want = P.sort_pairs(full.get_colliding_pairs())
        parts = [b.get_colliding_pairs() for b in ranks]
        assert all(len(p) > 0 for p in parts)
        got = P.sort_pairs(np.concatenate(parts))
        assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_morton_slab_trees(world):
    """The in-library decomposition (c2d_gpu_set_shard + c2d_gpu_set_slab_mode): every rank mirrors the
    shape state, owns one Morton-key slab of the Dynamic tree (+ the halo of higher slabs) and returns its share; the
    union over the ranks is the single-registry result, each pair once; the slabs are balanced although the scene is
    clustered; bullets and static shapes included."""
    sc = _scene()
    movable = sc["ent_body"] != 0
    full = H.Backend(SUBJECT)
    scenes.populate(full, sc)
    ranks = []
    for r in range(world):
        b = H.Backend(SUBJECT)
        scenes.populate(b, sc)
        capi.set_shard(b.native_context(), r, world)
        capi.set_slab_mode(b.native_context(), 1000)
        ranks.append(b)
    alive = np.ones(len(sc["ent_id"]), bool)
    for f, d in enumerate(_moves(sc, 4)):
        for b in [full] + ranks:
            b.move_translate(sc["ent_id"][movable], d[movable])
        if f == 2:  # a structural change in the middle: the slabs follow
            for b in [full] + ranks:
                b.remove_entities(sc["ent_id"][100:140])
            alive[100:140] = False
            movable = movable & alive
        want = P.sort_pairs(full.get_colliding_pairs())
        parts = [b.get_colliding_pairs() for b in ranks]
        stats = [capi.comm_stats(b.native_context()) for b in ranks]
        got = P.sort_pairs(np.concatenate(parts))
        assert len(want) > 5000
        assert got.tobytes() == want.tobytes()
        owned = np.array([s["slab_owned"] for s in stats])
        leaves = np.array([s["slab_leaves"] for s in stats])
        assert owned.sum() == int(((sc["ent_body"] == 1) & alive).sum())  # every Dynamic shape (one per entity) has one owner
        assert owned.min() > 0.8 * owned.sum() / world and owned.max() < 1.2 * owned.sum() / world, owned
        assert (leaves >= owned).all() and leaves.sum() < 1.6 * owned.sum(), (owned, leaves)


@pytest.mark.parametrize("world", [2, 4])
def test_weak_spatial_slabs_with_ghosts(world):
    sc = _scene()
    x = sc["ent_xf3"][:, 0]
    full = H.Backend(SUBJECT)
    scenes.populate(full, sc)
    ranks = []
    for r in range(world):
        owned, ghost = sharding.slab_selection(x, r, world, margin=40.0)
        local = np.nonzero(owned | ghost)[0]
        b = H.Backend(SUBJECT)
        ids = scenes.populate(b, {k: v[local] for k, v in sc.items()})
        b.set_ghost(ids[ghost[local]], np.ones(int(ghost[local].sum()), np.uint8))
        ranks.append((b, local))
        assert 0.8 * len(x) / world < owned.sum() < 1.2 * len(x) / world  # balanced despite the clustering
    movable = sc["ent_body"] != 0
    for d in _moves(sc):
        full.move_translate(sc["ent_id"][movable], d[movable])
        for b, local in ranks:
            m = movable[local]
            b.move_translate(sc["ent_id"][local][m], d[local][m])
        want = _entity_keys(full.get_colliding_pairs())
        got = np.sort(np.concatenate([_entity_keys(b.get_colliding_pairs()) for b, _ in ranks]))
        assert len(np.unique(got)) == len(got), "a pair was reported by two ranks"
        assert np.array_equal(want, got)
        assert len(want) > 5000


@pytest.mark.parametrize("world", [2, 5])
def test_ray_sharding(world):
    """c2d_gpu_set_ray_sharding: every rank is handed the same batch and casts its contiguous share; the other rays come
    back empty, and the per-ray lists of the ranks concatenate to the single-registry result (both tiers: a batch small
    enough for the latency path when unsharded, and a throughput batch)."""
    sc = _scene()
    full = H.Backend(SUBJECT)
    scenes.populate(full, sc)
    ranks = []
    for r in range(world):
        b = H.Backend(SUBJECT)
        scenes.populate(b, sc)
        capi.set_shard(b.native_context(), r, world)
        capi.set_ray_sharding(b.native_context(), True)
        ranks.append(b)
    for n in (40, 3000):
        ray4, inf = scenes.rays(n, 800.0, seed=3 + n, finite_len=70.0)
        woff, whits = full.raycast(ray4, inf)
        parts = []
        for r, b in enumerate(ranks):
            off, hits = b.raycast(ray4, inf)
            counts = (off[1:] - off[:-1]).astype(np.int64)
            lo, hi = (n * r) // world, (n * (r + 1)) // world
            assert counts[:lo].sum() == 0 and counts[hi:].sum() == 0 and int(off[-1]) == len(hits)
            parts.append((counts, hits))
        counts = sum(c for c, _ in parts)
        assert (counts == (woff[1:] - woff[:-1]).astype(np.int64)).all()
        assert np.concatenate([h for _, h in parts]).tobytes() == whits.tobytes()
