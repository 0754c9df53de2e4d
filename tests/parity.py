"""Parity helpers: canonical orientation, adjudication through the reference's own free
functions, and the record comparison used by every getCollidingPairs parity test.

TEST INFRASTRUCTURE.  `ref` is a tests.harness.Backend over oracle/_ref (the unmodified reference) or over
oracle/libc2d_oracle.so (the C restatement); `name` is that backend's library name for the free functions.
"""
from __future__ import annotations

import os

import numpy as np

from tests import harness as H

FLOAT_FIELDS = ("normal", "points")


class SceneIndex:
    """Per-shape lookup tables for a scene whose shapes were added in order (ShapeId i = i-th shape)."""

    def __init__(self, scene, shape_ids=None):
        self.scene = scene
        n = len(scene["shp_entity"])
        ids = np.arange(n, dtype=np.uint32) if shape_ids is None else np.asarray(shape_ids, np.uint32)
        size = int(ids.max()) + 1 if n else 0
        self.entity = np.zeros(size, np.uint32)
        self.stype = np.zeros(size, np.uint8)
        self.count = np.zeros(size, np.uint8)
        self.geom = np.zeros((size, 16), np.float32)
        self.entity[ids] = scene["shp_entity"]
        self.stype[ids] = scene["shp_type"]
        self.count[ids] = scene["shp_count"]
        self.geom[ids] = scene["shp_geom"]
        body_of_entity = dict(zip(scene["ent_id"].tolist(), scene["ent_body"].tolist()))
        self.body = np.array([body_of_entity[int(e)] for e in self.entity], np.uint8) if size else np.zeros(0, np.uint8)

    def batch(self, shapes):
        return H.ShapeBatch(self.stype[shapes], self.geom[shapes], self.count[shapes])


def pair_keys(p):
    """(entityA, shapeA, entityB, shapeB) tuples as a structured, sortable array."""
    k = np.zeros(len(p), dtype=[("eA", "<u4"), ("sA", "<u4"), ("eB", "<u4"), ("sB", "<u4")])
    k["eA"], k["sA"], k["eB"], k["sB"] = p["entityA"], p["shapeA"], p["entityB"], p["shapeB"]
    return k


def canonicalise(p, idx: SceneIndex):
    """Same-tree pairs (Dyn-Dyn, Bullet-Bullet) -> shapeA < shapeB; returns (records, swapped mask)."""
    out = p.copy()
    same = idx.body[p["shapeA"]] == idx.body[p["shapeB"]]
    swap = same & (p["shapeA"] > p["shapeB"])
    out["shapeA"][swap], out["shapeB"][swap] = p["shapeB"][swap], p["shapeA"][swap]
    out["entityA"][swap], out["entityB"][swap] = p["entityB"][swap], p["entityA"][swap]
    return out, swap


def adjudicate(name: str, ref: H.Backend, idx: SceneIndex, shapeA, shapeB):
    """Manifolds the reference's free functions give for (shapeA[i], shapeB[i]) in THAT orientation, using
    the registry's current (and, for bullets, previous) transforms: collide / sweep(1 moving) / sweep(2 moving)
    chosen like Registry::narrowPhaseCollision (Registry.hpp:626-703).  Returns (manifolds, hit flags)."""
    shapeA, shapeB = np.asarray(shapeA, np.int64), np.asarray(shapeB, np.int64)
    n = len(shapeA)
    man = np.zeros(n, H.MANIFOLD_DTYPE)
    hit = np.zeros(n, bool)
    if n == 0:
        return man, hit
    eA, eB = idx.entity[shapeA], idx.entity[shapeB]
    curA, curB = ref.get_transforms(eA), ref.get_transforms(eB)
    bA, bB = idx.body[shapeA] == H.BULLET, idx.body[shapeB] == H.BULLET
    A, B = idx.batch(shapeA), idx.batch(shapeB)
    m0 = ~bA & ~bB
    if m0.any():
        i = np.nonzero(m0)[0]
        mm, _ = H.collide(name, A.take(i), curA[i], B.take(i), curB[i])
        man[i] = mm
        hit[i] = mm["pointCount"] > 0
    m1 = bA & ~bB
    if m1.any():
        i = np.nonzero(m1)[0]
        prevA = ref.get_prev_transforms(eA[i])
        _, mm, hh = H.sweep(name, 1, A.take(i), prevA, curA[i], B.take(i), curB[i])
        man[i] = mm
        hit[i] = hh > 0
    m2 = bA & bB
    if m2.any():
        i = np.nonzero(m2)[0]
        prevA, prevB = ref.get_prev_transforms(eA[i]), ref.get_prev_transforms(eB[i])
        _, mm, hh = H.sweep(name, 2, A.take(i), prevA, curA[i], B.take(i), prevB, curB[i])
        man[i] = mm
        hit[i] = hh > 0
    assert not (~bA & bB).any(), "a lone bullet must be shape A (cross-tree pairs are bullet-first)"
    return man, hit


def expected_pairs(name: str, ref: H.Backend, idx: SceneIndex, ref_pairs, return_dropped=False):
    """Canonical expected records from the reference's getCollidingPairs output:
    swapped same-tree pairs get their manifold re-evaluated by the reference's own collide/sweep in canonical
    orientation.  Returns (records sorted by key, number of orientation-ambiguous pairs dropped)."""
    can, swap = canonicalise(ref_pairs, idx)
    dropped = 0
    dropped_records = can[:0]
    if swap.any():
        i = np.nonzero(swap)[0]
        man, hit = adjudicate(name, ref, idx, can["shapeA"][i], can["shapeB"][i])
        for f in ("normal", "points", "pointCount"):
            can[f][i] = man[f]
        keep = np.ones(len(can), bool)
        keep[i[~hit]] = False  # colliding only in the reference's own orientation (seg-seg noise, D8)
        dropped = int((~hit).sum())
        dropped_records = can[~keep]
        can = can[keep]
    order = np.argsort(pair_keys(can), order=("eA", "sA", "eB", "sB"), kind="stable")
    can = can[order]
    # PartitioningMethod::Grid reports a pair once per shared cell and may report (a,b) and (b,a):
    # the expected set is the UNIQUE canonical set
    keys = pair_keys(can)
    first = np.ones(len(can), bool)
    first[1:] = keys[1:] != keys[:-1]
    if return_dropped:
        return can[first], dropped_records
    return can[first], dropped


def sort_pairs(p):
    return p[np.argsort(pair_keys(p), order=("eA", "sA", "eB", "sB"))]


def manifold_close(a, b, rel=1e-5):
    """Per-record closeness of normals, points, anchors and separations within `rel` relative (north-star)."""
    if len(a) == 0:
        return np.ones(0, bool)
    ok = a["pointCount"] == b["pointCount"]
    fa = np.concatenate([a["normal"].reshape(len(a), -1), _pts(a)], axis=1)
    fb = np.concatenate([b["normal"].reshape(len(b), -1), _pts(b)], axis=1)
    tol = rel * np.maximum(1.0, np.maximum(np.abs(fa), np.abs(fb)))
    return ok & (np.abs(fa - fb) <= tol).all(axis=1)


def _pts(p):
    q = p["points"]
    return np.concatenate([q["point"].reshape(len(p), -1), q["anchorA"].reshape(len(p), -1),
                           q["anchorB"].reshape(len(p), -1), q["separation"].reshape(len(p), -1)], axis=1)


def manifold_bits_equal(a, b):
    if len(a) == 0:
        return np.ones(0, bool)
    fa = np.concatenate([a["normal"].reshape(len(a), -1), _pts(a)], axis=1).view(np.uint32)
    fb = np.concatenate([b["normal"].reshape(len(b), -1), _pts(b)], axis=1).view(np.uint32)
    # +0 / -0 compare equal
    return (a["pointCount"] == b["pointCount"]) & ((fa == fb) | (((fa | fb) & 0x7FFFFFFF) == 0)).all(axis=1)


def compare_pairs(name: str, ref: H.Backend, idx: SceneIndex, ref_pairs, got_pairs, max_ambiguous=None):
    """Full comparison.  Returns a dict of counts; raises AssertionError on any non-adjudicated difference."""
    exp, dropped = expected_pairs(name, ref, idx, ref_pairs)
    if os.environ.get("C2D_DRYRUN_CANON"):  # harness self-test: "got" comes from a non-canonical backend
        got_pairs, _ = expected_pairs(name, ref, idx, got_pairs)
    got = sort_pairs(got_pairs)
    ke, kg = pair_keys(exp), pair_keys(got)
    # same-tree records must come out canonical
    same = idx.body[got["shapeA"]] == idx.body[got["shapeB"]]
    assert (got["shapeA"][same] < got["shapeB"][same]).all(), "same-tree pair not emitted with shapeA < shapeB"
    assert len(np.unique(kg)) == len(kg), "duplicate pairs in the result"
    only_exp = np.setdiff1d(ke, kg)
    only_got = np.setdiff1d(kg, ke)
    ambiguous = dropped
    # a pair only the reference has must be one the reference itself rejects in canonical orientation;
    # those were dropped already, so anything left is a real miss.
    assert len(only_exp) == 0, f"{len(only_exp)} reference pairs missing, e.g. {only_exp[:5]}"
    if len(only_got):
        # pairs only we have: legitimate only if the reference's own narrowphase accepts them in canonical
        # orientation and rejects them in the reversed one (orientation noise, D8) on a same-tree pair.
        man, hit = adjudicate(name, ref, idx, only_got["sA"], only_got["sB"])
        assert hit.all(), f"{int((~hit).sum())} extra pairs the reference rejects, e.g. {only_got[~hit][:5]}"
        sameb = idx.body[only_got["sA"]] == idx.body[only_got["sB"]]
        assert sameb.all(), f"extra cross-tree pairs: {only_got[~sameb][:5]}"
        _, hit_rev = adjudicate(name, ref, idx, only_got["sB"], only_got["sA"])
        assert (~hit_rev).all(), "extra pair that the reference accepts in both orientations"
        ambiguous += len(only_got)
    if max_ambiguous is not None:
        assert ambiguous <= max_ambiguous, f"{ambiguous} orientation-ambiguous pairs (> {max_ambiguous})"
    common_e = exp[np.isin(ke, kg)]
    common_g = got[np.isin(kg, ke)]
    close = manifold_close(common_e, common_g)
    assert close.all(), (f"{int((~close).sum())} manifolds differ beyond 1e-5 rel, first: "
                         f"{common_e[~close][:1]} vs {common_g[~close][:1]}")
    bits = manifold_bits_equal(common_e, common_g)
    return dict(expected=len(exp), got=len(got), ambiguous=ambiguous, bit_identical=int(bits.sum()),
                not_bit_identical=int((~bits).sum()))
