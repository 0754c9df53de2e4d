"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

* free_functions_v1.npz — seeded random shape pairs of all 16 type pairs with the reference's collide manifolds and
  areColliding flags, one- and two-moving sweeps, shape raycasts, ray sweeps, computeAABB;
* registry_scene_v1.npz — a scripted registry scene (all body types, categories, bullets, transform moves,
  teleports, deactivation, removal) with, per frame, the reference's getCollidingPairs records in canonical
  orientation (tests/parity.py) and the rayCast hits of a ray batch.
The tests (tests/test_golden_fixtures.py) replay the same inputs on the oracle (CPU) and on the CUDA path (GPU).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from colli2de_b200 import scenes  # noqa: E402
from tests import gen, harness as H, parity as P  # noqa: E402


def scripted_scene():
    sc = scenes.mixed_scene(2500, 250, 220.0, seed=1234)
    sc["ent_body"][::33] = 2
    sc["shp_cat"][::6] = 2
    sc["shp_cat"][::10] = 3
    return sc


def script(frame, sc, rng):
    """The mutations applied before query `frame` (deterministic): list of (method name, args)."""
    n = len(sc["ent_id"])
    ids = np.arange(n, dtype=np.uint32)
    gone = ids[7::23]  # removed before query 3
    alive = np.ones(n, bool)
    if frame > 3:
        alive[gone] = False
    mov = ids[(sc["ent_body"] != 0) & alive]
    bul = ids[(sc["ent_body"] == 2) & alive]
    nonbul = ids[(sc["ent_body"] != 2) & alive]
    ops = [("move_translate", (mov, scenes.jitter(rng, len(mov), 0.6))),
           ("move_translate", (bul, scenes.jitter(rng, len(bul), 10.0)))]
    if frame % 2 == 1:
        d3 = np.zeros((len(nonbul), 3), np.float32)
        d3[:, :2] = scenes.jitter(rng, len(nonbul), 0.3)
        d3[:, 2] = rng.next_float(-0.25, 0.25, len(nonbul))
        ops.append(("move_transform", (nonbul, d3)))
    if frame == 2:
        tel = nonbul[::41]
        t3 = np.zeros((len(tel), 3), np.float32)
        t3[:, :2] = rng.next_vec2(0.0, 220.0, len(tel))
        t3[:, 2] = rng.next_float(0.0, 6.0, len(tel))
        ops.append(("teleport_transform", (tel, t3)))
        off = ids[4::19]
        ops.append(("set_active", (off, np.zeros(len(off), np.uint8))))
    if frame == 3:
        ops.append(("remove_entities", (gone,)))
    return ops


def main():
    assert H.available("reference"), "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"
    rs = np.random.default_rng(20260101)
    # ---- free functions
    ta, tb = gen.all_type_pairs(120)
    n = len(ta)
    A, B = gen.random_shapes(rs, n, ta), gen.random_shapes(rs, n, tb)
    xa, xb = H.xf5_from_xf3(gen.random_xf3(rs, n, 1.2)), H.xf5_from_xf3(gen.random_xf3(rs, n, 1.2))
    man, flag = H.collide("reference", A, xa, B, xb)
    a1 = xa.copy()
    a1[:, :2] += rs.uniform(-5, 5, (n, 2)).astype(np.float32)
    b1 = xb.copy()
    b1[:, :2] += rs.uniform(-3, 3, (n, 2)).astype(np.float32)
    s1 = H.sweep("reference", 1, A, xa, a1, B, xb)
    s2 = H.sweep("reference", 2, A, xa, a1, B, xb, b1)
    ray = np.zeros((n, 4), np.float32)
    ray[:, :2] = rs.uniform(-4, 4, (n, 2))
    inf = (np.arange(n) % 2).astype(np.uint8)
    tgt = rs.uniform(-1.5, 1.5, (n, 2))
    d = tgt - ray[:, :2]
    ray[:, 2:] = np.where(inf[:, None] == 1, d / np.linalg.norm(d, axis=1, keepdims=True), tgt + 0.5 * d)
    rt, rh = H.raycast_shape("reference", A, xa, ray, inf)
    st, sh = H.sweep_ray("reference", A, xa, a1, ray, inf)
    aabb = H.compute_aabb("reference", A, xa)
    np.savez_compressed(os.path.join(HERE, "free_functions_v1.npz"),
                        typeA=A.stype, geomA=A.geom, countA=A.count, typeB=B.stype, geomB=B.geom, countB=B.count,
                        xa=xa, xb=xb, xa_end=a1, xb_end=b1, manifold=man, colliding=flag,
                        sweep1_fraction=s1[0], sweep1_manifold=s1[1], sweep1_hit=s1[2],
                        sweep2_fraction=s2[0], sweep2_manifold=s2[1], sweep2_hit=s2[2],
                        ray=ray, infinite=inf, ray_times=rt, ray_hit=rh, raysweep_times=st, raysweep_hit=sh, aabb=aabb)
    # ---- registry scene
    sc = scripted_scene()
    idx = P.SceneIndex(sc)
    ref = H.Backend("reference")
    scenes.populate(ref, sc)
    rng = scenes.LCG(77)
    ray4, rinf = scenes.rays(400, 220.0, seed=8)
    masks = np.full(400, ~np.uint64(0))
    masks[::3] = 1
    out = {}
    for frame in range(5):
        for name, args in script(frame, sc, rng):
            getattr(ref, name)(*args)
        # pairs that collide only in the reference's own (tree-dependent) orientation are kept apart: an
        # implementation may or may not report them (segment-segment noise)
        exp, ambiguous = P.expected_pairs("reference", ref, idx, ref.get_colliding_pairs(), return_dropped=True)
        out[f"pairs_{frame}"] = exp
        out[f"ambiguous_{frame}"] = ambiguous
        off, hits = ref.raycast(ray4, rinf, masks)
        out[f"ray_offsets_{frame}"] = off
        out[f"ray_hits_{frame}"] = hits
    np.savez_compressed(os.path.join(HERE, "registry_scene_v1.npz"), ray4=ray4, ray_infinite=rinf, ray_masks=masks, **out)
    print("wrote fixtures:", {k: len(v) for k, v in out.items() if k.startswith("pairs")})


if __name__ == "__main__":
    main()
