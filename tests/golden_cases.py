"""Known-answer vectors restated from the reference's own Catch2 tests (which cannot be built here: no Catch2).
Each case cites the reference test it comes from.  run_* functions replay them on any rn_* backend
("oracle", "reference", "b200").  TEST INFRASTRUCTURE."""
from __future__ import annotations

import math

import numpy as np

from tests import harness as H

APPROX = 1e-3  # Approx(...).margin(0.001f) in the reference tests


# ---- shape builders ----------------------------------------------------------------------------------
def circle(cx, cy, r):
    return (H.CIRCLE, [cx, cy, r], 0)


def capsule(x1, y1, x2, y2, r):
    return (H.CAPSULE, [x1, y1, x2, y2, r], 0)


def segment(x1, y1, x2, y2):
    return (H.SEGMENT, [x1, y1, x2, y2], 0)


def rectangle(cx, cy, hw, hh, angle=0.0):
    """makeRectangle (reference src/geometry/Shapes.cpp:8-38)"""
    c, s = np.float32(math.cos(angle)), np.float32(math.sin(angle))
    g = []
    for lx, ly in ((-hw, -hh), (hw, -hh), (hw, hh), (-hw, hh)):
        lx, ly = np.float32(lx), np.float32(ly)
        g += [float(lx * c - ly * s + np.float32(cx)), float(lx * s + ly * c + np.float32(cy))]
    return (H.POLYGON, g, 4)


def polygon(*pts):
    g = []
    for x, y in pts:
        g += [x, y]
    return (H.POLYGON, g, len(pts))


def batch(shapes):
    n = len(shapes)
    geom = np.zeros((n, 16), np.float32)
    for i, (_, g, _) in enumerate(shapes):
        geom[i, :len(g)] = g
    return H.ShapeBatch([s[0] for s in shapes], geom, [s[2] for s in shapes])


def xf(tx=0.0, ty=0.0, angle=0.0):
    return [tx, ty, angle]


# ---- tests/collision/test_collision.cpp ------------------------------------------------------------------
# (name, A, xfA, B, xfB, expected) ; expected keys: count, normal (x, y, abs_x, abs_y), point, point1, sep, sep_lt, sep_le
T0 = xf()
RECT = rectangle(0, 0, 1, 1)
CAP = capsule(-1, 0, 1, 0, 0.5)
SEG = segment(-1, 0, 1, 0)
COLLIDE_CASES = [
    # Circle vs Circle  (test_collision.cpp:48-118)
    ("cc none", circle(0, 0, 1), T0, circle(3, 0, 1), T0, dict(count=0)),
    ("cc tangent right", circle(0, 0, 1), T0, circle(2, 0, 1), T0, dict(count=1, normal=(1, 0, True, False), point=(1, 0), sep=0)),
    ("cc tangent top", circle(0, 0, 1), T0, circle(0, 2, 1), T0, dict(count=1, normal=(0, 1, False, False), point=(0, 1), sep=0)),
    ("cc overlap", circle(0, 0, 1), T0, circle(1.5, 0, 1), T0, dict(count=1, point=(0.75, 0), sep=-0.5)),
    ("cc rotated", circle(1, 0, 1), xf(0, 0, 0.5), circle(1, 0, 1), xf(2, 0, 0.5),
     dict(count=1, normal=(1, 0, False, True), point=(math.cos(0.5) + 1.0, math.sin(0.5)), sep=0)),
    # Capsule vs Circle (test_collision.cpp:120-181)
    ("capc none", CAP, T0, circle(0, 3, 0.5), T0, dict(count=0)),
    ("capc tangent side", CAP, T0, circle(0, 1, 0.5), T0, dict(count=1, normal=(0, 1, False, False), point=(0, 0.5), sep=0)),
    ("capc tangent end", CAP, T0, circle(-1, -1, 0.5), T0, dict(count=1, normal=(0, -1, False, False), point=(-1, -0.5), sep=0)),
    ("capc deep", CAP, T0, circle(0, 0, 0.8), T0, dict(count=1, sep_lt=-0.5)),
    ("capc rotated", CAP, xf(0, 0, 0.5), circle(0, 1, 0.5), xf(0, 0, -0.5),
     dict(count=1, normal=(-0.479425, 0.877583, False, False), point=(0.608943, 0.640503), sep=-0.459698)),
    # Capsule vs Capsule (test_collision.cpp:183-245)
    ("capcap none", CAP, T0, capsule(4, 0, 6, 0, 0.5), T0, dict(count=0)),
    ("capcap side", CAP, T0, capsule(2, 0, 4, 0, 0.5), T0, dict(count=1, normal=(1, 0, True, False), point=(1.5, 0), sep=0)),
    ("capcap end", CAP, T0, capsule(0, 1, 0, 3, 0.5), T0, dict(count=1, normal=(0, 1, False, False), point=(0, 0.5), sep=0)),
    ("capcap cross", CAP, T0, capsule(0, -1, 0, 1, 0.5), T0, dict(count=1, sep_lt=0)),
    ("capcap rotated", CAP, xf(0, 0, 0.25), CAP, xf(0, 0, -0.25), dict(count=1, normal=(1, 0, False, True), point=(0, 0), sep=-1)),
    # Segment vs Segment (test_collision.cpp:247-303)
    ("ss none", SEG, T0, segment(3, 0, 5, 0), T0, dict(count=0)),
    ("ss endpoint", SEG, T0, segment(1, 0, 3, 0), T0, dict(count=1, point=(1, 0), sep=0)),
    ("ss crossing", SEG, T0, segment(0, -1, 0, 1), T0, dict(count=1, sep=0)),
    ("ss overlapping", SEG, T0, segment(-0.5, 0, 0.5, 0), T0, dict(count=1, sep=0)),
    ("ss rotated", SEG, xf(0, 0, 0.25), SEG, xf(0, 0, -0.25), dict(count=1, normal=(1, 0, False, True), point=(0, 0), sep=0)),
    # Polygon vs Circle (test_collision.cpp:305-369)
    ("pc none", RECT, T0, circle(0, 3, 0.5), T0, dict(count=0)),
    ("pc top", RECT, T0, circle(0, 1.5, 0.5), T0, dict(count=1, normal=(0, 1, False, False), point=(0, 1), sep=0)),
    ("pc right", RECT, T0, circle(1.5, 0, 0.5), T0, dict(count=1, normal=(1, 0, True, False), point=(1, 0), sep=0)),
    ("pc inside", RECT, T0, circle(0, 0, 0.5), T0, dict(count=1, sep_lt=0)),
    ("pc rotated", rectangle(0, 0, 1, 1, 0.3), xf(0, 0, 0.5), circle(1, 0, 0.5), xf(0, 0, 0.5),
     dict(count=1, normal=(-0.696704, -0.717353, False, False), point=(1.067317, 0.674784), sep=-0.544664)),
    # Circle vs Segment (test_collision.cpp:371-431)
    ("cs none", circle(0, 2, 0.5), T0, SEG, T0, dict(count=0)),
    ("cs center", circle(0, 0.5, 0.5), T0, SEG, T0, dict(count=1, normal=(0, -1, False, False), point=(0, 0), sep=0)),
    ("cs endpoint", circle(1.5, 0, 0.5), T0, SEG, T0, dict(count=1, point=(1, 0), sep=0)),
    ("cs overlap", circle(0, 0, 0.75), T0, SEG, T0, dict(count=1, sep_lt=0)),
    ("cs rotated", circle(0, 0.5, 0.5), xf(0, 0, 0.25), SEG, xf(0, 0, 0.25),
     dict(count=1, normal=(0.247404, -0.968912, False, False), point=(0, 0), sep=0)),
    # Capsule vs Polygon (test_collision.cpp:433-481): the capsule radius is ignored
    ("pcap none", RECT, T0, capsule(0, 3, 0, 4, 0.5), T0, dict(count=0)),
    ("pcap tangent top (radius ignored)", RECT, T0, capsule(0, 1.5, 0, 2.5, 0.5), T0, dict(count=0)),
    ("pcap tangent right (radius ignored)", RECT, T0, capsule(1.5, 0, 2.5, 0, 0.5), T0, dict(count=0)),
    ("pcap intersecting", RECT, T0, capsule(0, -1, 0, 1, 0.5), T0, dict(count=2, sep_le=0.01)),
    ("pcap rotated", rectangle(0, 0, 1, 1, 0.1), xf(0, 0, 0.3), capsule(0, -1, 0, 1, 0.5), xf(0, 0, 0.3), dict(count=2, sep_le=0.01)),
    # Capsule vs Segment (test_collision.cpp:483-529)
    ("caps none", CAP, T0, segment(3, -2, 3, 2), T0, dict(count=0)),
    ("caps side", CAP, T0, segment(1.5, -2, 1.5, 2), T0, dict(count=1)),
    ("caps end", CAP, T0, segment(0, 0.5, 0, 2.5), T0, dict(count=1)),
    ("caps overlap", CAP, T0, segment(0, -2, 0, 2), T0, dict(count=1, sep_lt=0)),
    ("caps rotated", CAP, xf(0, 0, 0.25), SEG, xf(0, 0, 0.25),
     dict(count=1, normal=(1, 0, False, True), point=(-0.718912, -0.247404), sep=-0.5)),
    # Polygon vs Segment (test_collision.cpp:531-580)
    ("ps none", RECT, T0, segment(-2, 3, 2, 3), T0, dict(count=0)),
    ("ps top", RECT, T0, segment(-2, 1, 2, 1), T0, dict(count=2)),
    ("ps right", RECT, T0, segment(1, -2, 1, 2), T0, dict(count=2, normal=(1, 0, True, False))),
    ("ps cross", RECT, T0, segment(0, -2, 0, 2), T0, dict(count=2, sep_le=0.01)),
    ("ps rotated", rectangle(0, 0, 1, 1, 0.2), xf(0, 0, 0.2), SEG, xf(0, 0, 0.2), dict(count=2, sep_le=0.01)),
    # Polygon vs Polygon (test_collision.cpp:582-650)
    ("pp none", RECT, T0, rectangle(3, 0, 1, 1), T0, dict(count=0)),
    ("pp right", RECT, T0, rectangle(1.999, 0, 1, 1), T0, dict(count=2, normal=(1, 0, False, True), sep_le=0.01)),
    ("pp top", RECT, T0, rectangle(0, 1.999, 1, 1), T0, dict(count=2, normal=(0, 1, True, False), sep_le=0.01)),
    ("pp overlap", RECT, T0, rectangle(0.5, 0.5, 1, 1), T0, dict(count=2, sep_lt=0)),
    ("pp rotated overlap", rectangle(0, 0, 1, 1, 0.5), T0, rectangle(0.5, 0.5, 1, 1, 0.25), T0,
     dict(count=2, normal=(0.877579, 0.479423, False, False), point=(-0.610249, 0.806114), point1=(-0.221508, -0.716316),
          sep=-1.537806)),
]


def run_collide_cases(name):
    """Replays COLLIDE_CASES through collide()/areColliding() incl. the A<->B symmetry of collideSymmetric
    (test_collision.cpp:14-46).  Returns the number of cases checked."""
    A = batch([c[1] for c in COLLIDE_CASES])
    B = batch([c[3] for c in COLLIDE_CASES])
    xa = H.xf5_from_xf3(np.float32([c[2] for c in COLLIDE_CASES]))
    xb = H.xf5_from_xf3(np.float32([c[4] for c in COLLIDE_CASES]))
    m, f = H.collide(name, A, xa, B, xb)
    mr, fr = H.collide(name, B, xb, A, xa)
    for i, case in enumerate(COLLIDE_CASES):
        label, exp = case[0], case[5]
        assert m["pointCount"][i] == exp["count"], (label, m[i])
        assert bool(f[i]) == (exp["count"] > 0), label
        # symmetry: reverse(collide(B, A)) == collide(A, B)
        assert mr["pointCount"][i] == m["pointCount"][i], label
        assert bool(fr[i]) == bool(f[i]), label
        assert np.allclose(np.abs(m["normal"][i]), np.abs(mr["normal"][i]), atol=1e-6), label
        for k in range(exp["count"]):
            assert np.allclose(m["points"]["point"][i, k], mr["points"]["point"][i, k], atol=1e-6), label
            assert np.allclose(m["points"]["anchorA"][i, k], mr["points"]["anchorB"][i, k], atol=1e-6), label
            assert abs(m["points"]["separation"][i, k] - mr["points"]["separation"][i, k]) < 1e-4, label
        if "normal" in exp:
            nx, ny, ax, ay = exp["normal"]
            gx, gy = m["normal"][i]
            assert abs((abs(gx) if ax else gx) - nx) < APPROX and abs((abs(gy) if ay else gy) - ny) < APPROX, (label, gx, gy)
        if "point" in exp:
            assert np.allclose(m["points"]["point"][i, 0], exp["point"], atol=APPROX), (label, m["points"]["point"][i, 0])
        if "point1" in exp:
            assert np.allclose(m["points"]["point"][i, 1], exp["point1"], atol=APPROX), label
        for k in range(exp["count"]):
            s = float(m["points"]["separation"][i, k])
            if "sep" in exp and (k == 0 or "point1" in exp):
                assert abs(s - exp["sep"]) < (0.01 if label == "cs endpoint" else APPROX), (label, s)
            if "sep_lt" in exp:
                assert s < exp["sep_lt"], (label, s)
            if "sep_le" in exp:
                assert s <= exp["sep_le"], (label, s)
    return len(COLLIDE_CASES)


# ---- tests/geometry/test_SweepShapes.cpp ------------------------------------------------------------------
# (name, moving, start, end, target, expected) — target transform is identity
SWEEP_CASES = [
    ("hits stationary circle :9-23", circle(-2, 0, 1), T0, xf(4, 0, 0), circle(1, 0, 1), dict(hit=True, fraction=0.25, normal=(1, 0), point=(0, 0), sep=0)),
    ("starting overlapped :25-38", circle(0, 0, 1), T0, xf(2, 0, 0), circle(1, 0, 1), dict(hit=True, fraction=0.0, normal=(1, 0), sep_lt=0)),
    ("misses :40-47", circle(-2, 0, 1), T0, xf(-3, 0, 0), circle(1, 0, 1), dict(hit=False)),
    ("vertical :49-63", circle(0, -2, 1), T0, xf(0, 4, 0), circle(0, 1, 1), dict(hit=True, fraction=0.25, normal=(0, 1), point=(0, 0), sep=0)),
    ("diagonal :65-78", circle(-3, -3, 1), T0, xf(4, 4, 0), circle(0, 0, 1), dict(hit=True, fraction=0.3964, normal=(math.sqrt(0.5), math.sqrt(0.5)), sep=0)),
    ("zero translation separated :80-87", circle(-2, 0, 1), T0, T0, circle(1, 0, 1), dict(hit=False)),
    ("rotates while moving :89-102", circle(-2, 0, 1), T0, xf(4, 0, math.pi), circle(1, 0, 1), dict(hit=True, fraction=0.25, normal=(math.sqrt(0.5), math.sqrt(0.5)))),
    ("zero translation overlapping :104-117", circle(0, 0, 1), T0, T0, circle(1, 0, 1), dict(hit=True, fraction=0.0, normal=(1, 0), sep_lt=0)),
    ("circle hits capsule :119-133", circle(0, 0, 0.5), T0, xf(2, 0, 0), capsule(2, -1, 2, 1, 0.5), dict(hit=True, fraction=0.5, normal=(1, 0), point=(1.5, 0), sep=0)),
    ("capsule hits circle :135-149", capsule(-2, -1, -2, 1, 0.5), T0, xf(4, 0, 0), circle(1, 0, 0.5), dict(hit=True, fraction=0.5, normal=(1, 0), point=(0.5, 0), sep=0)),
    ("circle hits rectangle :151-168", circle(-2, 0, 0.5), T0, xf(3, 0, 0), RECT, dict(hit=True, fraction=1.0 / 6.0, normal=(1, 0), point=(-1, 0), sep=0)),
]


def run_sweep_cases(name):
    A = batch([c[1] for c in SWEEP_CASES])
    B = batch([c[4] for c in SWEEP_CASES])
    s0 = H.xf5_from_xf3(np.float32([c[2] for c in SWEEP_CASES]))
    s1 = H.xf5_from_xf3(np.float32([c[3] for c in SWEEP_CASES]))
    tb = H.xf5_from_xf3(np.float32([T0 for _ in SWEEP_CASES]))
    frac, man, hit = H.sweep(name, 1, A, s0, s1, B, tb)
    for i, case in enumerate(SWEEP_CASES):
        label, exp = case[0], case[5]
        assert bool(hit[i]) == exp["hit"], label
        if not exp["hit"]:
            continue
        assert abs(float(frac[i]) - exp["fraction"]) < 1e-3, (label, frac[i])
        assert man["pointCount"][i] == 1, label
        assert np.allclose(man["normal"][i], exp["normal"], atol=APPROX), (label, man["normal"][i])
        if "point" in exp:
            assert np.allclose(man["points"]["point"][i, 0], exp["point"], atol=APPROX), (label, man["points"]["point"][i, 0])
        s = float(man["points"]["separation"][i, 0])
        if "sep" in exp:
            assert abs(s - exp["sep"]) < APPROX, (label, s)
        if "sep_lt" in exp:
            assert s < exp["sep_lt"], label
    return len(SWEEP_CASES)


# ---- tests/geometry/test_RaycastShapes.cpp, test_SweepRaycast.cpp -------------------------------------------
# (name, shape, ray4, infinite, expected (entry, exit) or None)
RAY_CASES = [
    ("circle finite through :10-18", circle(1, 1, 1), (-2, 1, 4, 1), 0, (1 / 3, 2 / 3)),
    ("circle starting inside :20-28", circle(0, 0, 1), (0, 0, 2, 0), 0, (0.0, 0.5)),
    ("circle infinite miss :30-35", circle(0, 0, 1), (0, 2, 1, 0), 1, None),
    ("segment finite hit :37-44", segment(1, 1, 3, 1), (2, 0, 2, 2), 0, (0.5, 0.5)),
    ("segment parallel miss :46-51", segment(0, 1, 3, 1), (0, 2, 3, 2), 0, None),
    ("polygon through square :53-61", rectangle(1, 1, 1, 1), (-1, 1, 3, 1), 0, (0.25, 0.75)),
    ("polygon infinite inside :63-71", polygon((0, 0), (2, 0), (1, 2)), (1, 1, 0, 1), 1, (0.0, "positive")),
    ("capsule through body :73-81", capsule(0, 0, 4, 0, 1), (-2, 0, 6, 0), 0, (0.125, 0.875)),
    ("capsule infinite miss :83-88", capsule(0, 0, 4, 0, 1), (0, 3, 1, 0), 1, None),
]
# (name, shape, start, end, ray4, infinite, expected)
RAY_SWEEP_CASES = [
    ("circle sweep vs ray :10-19", circle(0, 0, 1), xf(-2, 0, 0), xf(2, 0, 0), (0, -2, 0, 2), 0, (0.5, 0.5)),
    ("circle sweep vs ray miss :21-28", circle(0, 0, 1), xf(-2, 4, 0), xf(2, 4, 0), (0, -2, 0, 2), 0, None),
    ("capsule sweep vs infinite ray :30-43", capsule(0, -1, 0, 1, 0.5), xf(-2, 0, 0), xf(2, 0, 0), (0, -2, 0, 1), 1, "ordered"),
]


def run_ray_cases(name):
    S = batch([c[1] for c in RAY_CASES])
    t, h = H.raycast_shape(name, S, H.xf5_from_xf3(np.float32([T0] * len(RAY_CASES))),
                           np.float32([c[2] for c in RAY_CASES]), np.uint8([c[3] for c in RAY_CASES]))
    for i, (label, _, _, _, exp) in enumerate(RAY_CASES):
        assert bool(h[i]) == (exp is not None), label
        if exp is None:
            continue
        assert abs(float(t[i, 0]) - exp[0]) < 1e-4, (label, t[i])
        if exp[1] == "positive":
            assert t[i, 1] > 0, label
        else:
            assert abs(float(t[i, 1]) - exp[1]) < 1e-4, (label, t[i])
    S = batch([c[1] for c in RAY_SWEEP_CASES])
    t, h = H.sweep_ray(name, S, H.xf5_from_xf3(np.float32([c[2] for c in RAY_SWEEP_CASES])),
                       H.xf5_from_xf3(np.float32([c[3] for c in RAY_SWEEP_CASES])),
                       np.float32([c[4] for c in RAY_SWEEP_CASES]), np.uint8([c[5] for c in RAY_SWEEP_CASES]))
    for i, (label, _, _, _, _, _, exp) in enumerate(RAY_SWEEP_CASES):
        assert bool(h[i]) == (exp is not None), label
        if exp is None:
            continue
        if exp == "ordered":
            assert t[i, 1] >= t[i, 0], label
        else:
            assert abs(float(t[i, 0]) - exp[0]) < 1e-4 and abs(float(t[i, 1]) - exp[1]) < 1e-4, (label, t[i])
    return len(RAY_CASES) + len(RAY_SWEEP_CASES)


# ---- tests/registry/test_registry_basic.cpp, test_registry_raycast.cpp ------------------------------------------
UNIT = np.float32([[0, 0, 1] + [0] * 13])


def _entity_pairs(p):
    return sorted((min(int(a), int(b)), max(int(a), int(b))) for a, b in zip(p["entityA"], p["entityB"]))


def _add_circle(b, ent, cat=None, mask=None):
    return int(b.add_shapes([ent], [H.CIRCLE], UNIT, cat=cat, mask=mask)[0])


def run_registry_cases(name):
    """test_registry_basic.cpp: pair sets :13-89, shape-id reuse :110-124, bullets :142-153, all body types
    :182-226, rectangle approach :228-253, category filter :156-168 (getCollidingPairs part), static-static
    :313-323, bullet-vs-bullet manifold :336-370."""
    # --- :13-89
    b = H.Backend(name)
    b.create_entities([1, 2], [H.STATIC, H.DYNAMIC], [[0, 0, 0], [3, 0, 0]])
    s1, s2 = _add_circle(b, 1), _add_circle(b, 2)
    assert _entity_pairs(b.get_colliding_pairs()) == []
    b.teleport_transform([2], [[1, 0, 0]])
    assert _entity_pairs(b.get_colliding_pairs()) == [(1, 2)]
    b.create_entities([3, 4, 5], [H.DYNAMIC] * 3, [[2, 0, 0], [1, 2, 0], [1, 5, 0]])
    for e in (3, 4, 5):
        _add_circle(b, e)
    assert _entity_pairs(b.get_colliding_pairs()) == [(1, 2), (1, 3), (2, 3), (2, 4)]
    b.set_active([s2], [0])
    assert _entity_pairs(b.get_colliding_pairs()) == [(1, 3)]
    b.set_active([s2], [1])
    assert len(b.get_colliding_pairs()) == 4
    b.remove_entities([2])
    assert _entity_pairs(b.get_colliding_pairs()) == [(1, 3)]
    b.remove_entities([1])
    assert len(b.get_colliding_pairs()) == 0
    assert b.size() == 3
    # --- shape id reuse :110-124
    b = H.Backend(name)
    b.create_entities([1, 2], [H.STATIC, H.STATIC], [[0, 0, 0], [5, 0, 0]])
    id1, id2 = _add_circle(b, 1), _add_circle(b, 2)
    b.remove_entities([2])
    b.create_entities([3], [H.STATIC], [[10, 0, 0]])
    id3 = _add_circle(b, 3)
    assert id3 == id2 and id1 != id3
    # --- category filter :129-136 (mask ignored by getCollidingPairs)
    b = H.Backend(name)
    b.create_entities([1, 2], [H.DYNAMIC, H.DYNAMIC], [[0, 0, 0], [0.5, 0, 0]])
    _add_circle(b, 1, [1], [1])
    _add_circle(b, 2, [2], [2])
    assert len(b.get_colliding_pairs()) == 0
    # --- all body types :182-226
    b = H.Backend(name)
    b.create_entities([1, 2, 3], [H.STATIC, H.DYNAMIC, H.BULLET], [[0, 0, 0], [0.5, 0, 0], [1, 0, 0]])
    for e in (1, 2, 3):
        _add_circle(b, e)
    assert _entity_pairs(b.get_colliding_pairs()) == [(1, 2), (1, 3), (2, 3)]
    b.teleport_translation([3], [[-3, 0]])
    assert len(b.get_colliding_pairs()) == 3  # the bullet's sweep still crosses both
    b.teleport_translation([2], [[3, 0]])
    b.move_translate([3], [[0, 0]])
    assert len(b.get_colliding_pairs()) == 0
    b.teleport_translation([2], [[2.05, 0]])
    assert len(b.get_colliding_pairs()) == 0
    b.teleport_translation([2], [[1.95, 0]])
    assert _entity_pairs(b.get_colliding_pairs()) == [(1, 2)]
    # --- rectangle approach :228-253
    b = H.Backend(name)
    b.create_entities([1, 2], [H.STATIC, H.DYNAMIC], [[0, 0, 0], [0, 19.5, 0]])
    big, small = rectangle(0, 0, 10, 5), rectangle(0, 0, 2, 5)
    for ent, r in ((1, big), (2, small)):
        g = np.zeros((1, 16), np.float32)
        g[0, :8] = r[1]
        b.add_shapes([ent], [H.POLYGON], g, [4])
    assert len(b.get_colliding_pairs()) == 0
    for _ in range(9):
        b.move_transform([2], [[0, -1, 0]])
        assert len(b.get_colliding_pairs()) == 0
    b.move_transform([2], [[0, -1, 0]])
    assert len(b.get_colliding_pairs()) == 1
    # --- static entities never collide :313-323
    b = H.Backend(name)
    b.create_entities([1, 2], [H.STATIC, H.STATIC], [[0, 0, 0], [0.5, 0, 0]])
    _add_circle(b, 1)
    _add_circle(b, 2)
    assert len(b.get_colliding_pairs()) == 0
    # --- bullet vs bullet switching places :336-370: manifold point (2.5, 0), |normal.x| = 1
    b = H.Backend(name)
    b.create_entities([1, 2], [H.BULLET, H.BULLET], [[0, 0, 0], [5, 0, 0]])
    _add_circle(b, 1)
    _add_circle(b, 2)
    b.teleport_translation([1], [[5, 0]])
    b.teleport_translation([2], [[0, 0]])
    p = b.get_colliding_pairs()
    assert len(p) == 1 and _entity_pairs(p) == [(1, 2)]
    assert p["pointCount"][0] == 1
    assert np.allclose(p["points"]["point"][0, 0], (2.5, 0.0), atol=1e-4)
    assert abs(abs(float(p["normal"][0, 0])) - 1.0) < 1e-4 and abs(float(p["normal"][0, 1])) < 1e-4
    return 8


def run_registry_ray_cases(name):
    """test_registry_basic.cpp:256-311 (finite ray with mask, infinite ray hits both) and D10 (coincident hits)."""
    b = H.Backend(name)
    b.create_entities([1, 2], [H.STATIC, H.STATIC], [[0, 0, 0], [4, 0, 0]])
    _add_circle(b, 1, [1], [1])
    _add_circle(b, 2, [2], [2])
    off, hits = b.raycast([[-2, 0, 5, 0]], [0], [1])
    assert len(hits) == 1 and hits["entity"][0] == 1
    assert abs(hits["entry"][0, 0] + 1.0) < 1e-5 and abs(hits["exit"][0, 0] - 1.0) < 1e-5
    assert abs(hits["entryTime"][0] - 1.0 / 7.0) < 1e-5 and abs(hits["exitTime"][0] - 3.0 / 7.0) < 1e-5
    b = H.Backend(name)
    b.create_entities([1, 2], [H.STATIC, H.STATIC], [[0, 0, 0], [4, 0, 0]])
    _add_circle(b, 1)
    _add_circle(b, 2)
    off, hits = b.raycast([[-1, 0, 1, 0]], [1])
    assert list(hits["entity"]) == [1, 2]
    assert np.allclose(hits["entryTime"], [0, 4], atol=1e-5) and np.allclose(hits["exitTime"], [2, 6], atol=1e-5)
    assert np.allclose(hits["entry"][:, 0], [-1, 3], atol=1e-5) and np.allclose(hits["exit"][:, 0], [1, 5], atol=1e-5)
    # SURVEY.md D10 [probe]: two coincident unit circles on different entities + one ray -> ONE hit
    b = H.Backend(name)
    b.create_entities([1, 2], [H.STATIC, H.STATIC], [[0, 0, 0], [0, 0, 0]])
    _add_circle(b, 1)
    _add_circle(b, 2)
    off, hits = b.raycast([[-3, 0, 3, 0]], [0])
    assert len(hits) == 1
    return 3
