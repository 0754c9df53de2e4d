"""Synthetic scene generator for the BASELINE.json configurations.

All randomness comes from the reference's test LCG (tests/utils/Random.hpp:14-34):
``state = 1664525*state + 1013904223``; ``next() = (state & 0xFFFFFF) / 2**24``;
``nextFloat(a, b) = a + (b - a) * next()`` in binary32.  The stream is produced with a
log-step jump-ahead so a 1 M-shape scene takes milliseconds, not a Python loop.

A scene is a plain dict of numpy arrays that every backend (reference, oracle, B200)
consumes unchanged:

    ent_id (E,) u4     ent_body (E,) u1 (0 Static, 1 Dynamic, 2 Bullet)     ent_xf3 (E,3) f4
    shp_entity (S,) u4 shp_type (S,) u1 (0 circle 1 capsule 2 segment 3 polygon)
    shp_geom (S,16) f4 shp_count (S,) u1   shp_cat (S,) u8   shp_mask (S,) u8
"""
from __future__ import annotations

import numpy as np

_A = np.uint32(1664525)
_C = np.uint32(1013904223)


class LCG:
    """Vectorised DeterministicRNG (tests/utils/Random.hpp:14-34)."""

    def __init__(self, seed: int = 12345):
        self.state = np.uint32(seed)

    def _states(self, n: int) -> np.ndarray:
        out = np.empty(n, np.uint32)
        if n == 0:
            return out
        with np.errstate(over="ignore"):
            out[0] = _A * self.state + _C
            m, a, c = 1, _A, _C
            while m < n:
                k = min(m, n - m)
                out[m:m + k] = a * out[:k] + c
                a, c = np.uint32(a * a), np.uint32(a * c + c)
                m *= 2
        self.state = out[-1]
        return out

    def next(self, n: int) -> np.ndarray:
        s = self._states(n)
        return (s & np.uint32(0x00FFFFFF)).astype(np.float32) / np.float32(0x01000000)

    def next_float(self, lo: float, hi: float, n: int) -> np.ndarray:
        lo, hi = np.float32(lo), np.float32(hi)
        return lo + (hi - lo) * self.next(n)

    def next_vec2(self, lo: float, hi: float, n: int) -> np.ndarray:
        """n pairs drawn x0,y0,x1,y1,... like generateRandomCircles (Random.hpp:62-76)."""
        v = self.next_float(lo, hi, 2 * n)
        return v.reshape(n, 2).copy()


def regular_polygon(nverts: np.ndarray, radius: float, n: int) -> np.ndarray:
    """Local-space regular polygons like makeRegularPolygon (Shapes.cpp:48-62), centre (0,0), rotation 0."""
    geom = np.zeros((n, 16), np.float32)
    for k in range(8):
        step = (np.float32(2.0) * np.float32(np.pi)) / nverts.astype(np.float32)
        theta = (step * np.float32(k)).astype(np.float32)
        use = k < nverts
        geom[:, 2 * k] = np.where(use, np.float32(radius) * np.cos(theta).astype(np.float32), 0)
        geom[:, 2 * k + 1] = np.where(use, np.float32(radius) * np.sin(theta).astype(np.float32), 0)
    return geom


def _mixed_geom(idx: np.ndarray):
    """cfg 2/3 mix by i%4: circle r=1 / capsule (+-1,0) r=.5 / segment (+-1.5,0) / regular polygon 3+i%6 verts R=1.5."""
    n = len(idx)
    stype = (idx % 4).astype(np.uint8)
    geom = np.zeros((n, 16), np.float32)
    count = np.zeros(n, np.uint8)
    c = stype == 0
    geom[c, 2] = 1.0
    k = stype == 1
    geom[k, 0], geom[k, 2], geom[k, 4] = -1.0, 1.0, 0.5
    s = stype == 2
    geom[s, 0], geom[s, 2] = -1.5, 1.5
    p = stype == 3
    nv = (3 + idx % 6).astype(np.int64)
    poly = regular_polygon(nv, 1.5, n)
    geom[p] = poly[p]
    count[p] = nv[p]
    return stype, geom, count


def _assemble(ent_body, ent_xf3, stype, geom, count, cat=None, mask=None):
    E = len(ent_body)
    return dict(
        ent_id=np.arange(E, dtype=np.uint32),
        ent_body=np.asarray(ent_body, np.uint8),
        ent_xf3=np.asarray(ent_xf3, np.float32).reshape(E, 3),
        shp_entity=np.arange(E, dtype=np.uint32),
        shp_type=np.asarray(stype, np.uint8),
        shp_geom=np.asarray(geom, np.float32).reshape(E, 16),
        shp_count=np.asarray(count, np.uint8),
        shp_cat=np.ones(E, np.uint64) if cat is None else np.asarray(cat, np.uint64),
        shp_mask=np.full(E, ~np.uint64(0)) if mask is None else np.asarray(mask, np.uint64),
    )


def repo_perf_scene(n: int, seed: int = 42, all_bullets: bool = False):
    """cfg 1a — tests/registry/test_registry_performance.cpp:114-162 (and 164-204 for bullets):
    one circle r=12 per entity at a random centre in [-1920, 3840]^2; Bullet if i%1000==0, Static if
    i%3==0, else Dynamic; every entity is then moved by Translation(3,3) (returned as 'first_move')."""
    rng = LCG(seed)
    centres = rng.next_vec2(-1920.0, 3840.0, n)
    i = np.arange(n)
    if all_bullets:
        body = np.full(n, 2, np.uint8)
    else:
        bullet = (i % 1000) == 0
        static = (~bullet) & ((i % 3) == 0)
        body = np.where(static, 0, np.where(bullet, 2, 1)).astype(np.uint8)
    xf3 = np.zeros((n, 3), np.float32)
    xf3[:, :2] = centres
    geom = np.zeros((n, 16), np.float32)
    geom[:, 2] = 12.0
    sc = _assemble(body, xf3, np.zeros(n, np.uint8), geom, np.zeros(n, np.uint8))
    sc["first_move"] = np.tile(np.float32([3.0, 3.0]), (n, 1))
    return sc


def baseline_cfg1(seed: int = 12345, n_dyn: int = 10_000, n_static: int = 1_000, box: float = 500.0):
    """cfg 1b — 10 k Dynamic circles r=2 + 1 k Static regular polygons (3-8 verts, R=3, random angle)."""
    rng = LCG(seed)
    n = n_dyn + n_static
    pos = rng.next_vec2(0.0, box, n)
    ang = rng.next_float(0.0, 2.0 * np.pi, n)
    i = np.arange(n)
    body = np.where(i < n_dyn, 1, 0).astype(np.uint8)
    stype = np.where(i < n_dyn, 0, 3).astype(np.uint8)
    geom = np.zeros((n, 16), np.float32)
    geom[:n_dyn, 2] = 2.0
    nv = (3 + i % 6).astype(np.int64)
    poly = regular_polygon(nv, 3.0, n)
    geom[n_dyn:] = poly[n_dyn:]
    count = np.where(i < n_dyn, 0, nv).astype(np.uint8)
    xf3 = np.zeros((n, 3), np.float32)
    xf3[:, :2] = pos
    xf3[n_dyn:, 2] = ang[n_dyn:]
    return _assemble(body, xf3, stype, geom, count)


def mixed_scene(n_dyn: int, n_static: int, box: float, seed: int = 4242, clustered: bool = False,
                static_radius: float = 2.0):
    """cfg 2 (100 k + 10 k in [0,1600]^2) and cfg 3 (1 M + 100 k in [0,5000]^2).

    Dynamic i: type i%4 (see _mixed_geom), random angle.  Static: regular polygons R=2.
    clustered=True: power-law cluster populations (rank^-1.2), Gaussian radius ~ sqrt(population),
    same mean density (SURVEY.md §8(d) cfg 3)."""
    rng = LCG(seed)
    n = n_dyn + n_static
    if clustered:
        n_clusters = max(8, n // 2000)
        centres = rng.next_vec2(0.0, box, n_clusters)
        w = (np.arange(1, n_clusters + 1, dtype=np.float64)) ** -1.2
        w /= w.sum()
        pop = np.maximum(1, np.floor(w * n)).astype(np.int64)
        owner = np.repeat(np.arange(n_clusters), pop)
        if len(owner) < n:
            owner = np.concatenate([owner, np.zeros(n - len(owner), np.int64)])
        owner = owner[:n]
        # interleave so every body type sees every cluster
        perm = (np.arange(n, dtype=np.int64) * 7919) % n
        owner = owner[perm]
        # Box-Muller from the LCG stream
        u1 = np.maximum(rng.next(n), np.float32(1e-7)).astype(np.float64)
        u2 = rng.next(n).astype(np.float64)
        r = np.sqrt(-2.0 * np.log(u1))
        gx, gy = r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)
        # mean density n/box^2  => cluster sigma so that area ~ pop / density
        dens = n / float(box * box)
        sigma = np.sqrt(pop[owner] / dens / (2 * np.pi))
        pos = np.empty((n, 2), np.float32)
        pos[:, 0] = (centres[owner, 0] + sigma * gx).astype(np.float32)
        pos[:, 1] = (centres[owner, 1] + sigma * gy).astype(np.float32)
    else:
        pos = rng.next_vec2(0.0, box, n)
    ang = rng.next_float(0.0, 2.0 * np.pi, n)
    i = np.arange(n)
    body = np.where(i < n_dyn, 1, 0).astype(np.uint8)
    stype, geom, count = _mixed_geom(i)
    nv = (3 + i % 6).astype(np.int64)
    spoly = regular_polygon(nv, static_radius, n)
    st = i >= n_dyn
    stype[st] = 3
    geom[st] = spoly[st]
    count[st] = nv[st]
    xf3 = np.zeros((n, 3), np.float32)
    xf3[:, :2] = pos
    xf3[:, 2] = ang
    return _assemble(body, xf3, stype, geom, count)


def bullet_scene(n_static: int = 1_000_000, n_bullet: int = 10_000, box: float = 3200.0, seed: int = 99):
    """cfg 4 — static polygons (R=1, 3-8 verts) + bullet circles r in {0.3, 0.5}."""
    rng = LCG(seed)
    n = n_static + n_bullet
    pos = rng.next_vec2(0.0, box, n)
    ang = rng.next_float(0.0, 2.0 * np.pi, n)
    i = np.arange(n)
    st = i < n_static
    body = np.where(st, 0, 2).astype(np.uint8)
    stype = np.where(st, 3, 0).astype(np.uint8)
    nv = (3 + i % 6).astype(np.int64)
    geom = regular_polygon(nv, 1.0, n)
    geom[~st] = 0
    geom[~st, 2] = np.where((i[~st] % 2) == 0, np.float32(0.3), np.float32(0.5))
    count = np.where(st, nv, 0).astype(np.uint8)
    xf3 = np.zeros((n, 3), np.float32)
    xf3[:, :2] = pos
    xf3[st, 2] = ang[st]
    return _assemble(body, xf3, stype, geom, count)


def jitter(rng: LCG, n: int, amp: float = 0.5) -> np.ndarray:
    """Per-frame move U(-amp, amp)^2."""
    return rng.next_vec2(-amp, amp, n)


def rays(n: int, box: float, seed: int = 7, finite_len: float = 50.0):
    """cfg 5 — odd index: finite ray of length 50 in a random direction; even index: infinite ray, unit dir."""
    rng = LCG(seed)
    start = rng.next_vec2(0.0, box, n)
    ang = rng.next_float(0.0, 2.0 * np.pi, n)
    d = np.stack([np.cos(ang).astype(np.float32), np.sin(ang).astype(np.float32)], axis=1)
    infinite = ((np.arange(n) % 2) == 0).astype(np.uint8)
    ray4 = np.zeros((n, 4), np.float32)
    ray4[:, :2] = start
    ray4[:, 2:] = np.where(infinite[:, None] == 1, d, start + np.float32(finite_len) * d)
    return ray4, infinite


def populate(backend, scene, chunk: int = 1 << 20):
    """Replay a scene dict onto a tests.harness.Backend-like object; returns the ShapeIds."""
    backend.create_entities(scene["ent_id"], scene["ent_body"], scene["ent_xf3"])
    return backend.add_shapes(scene["shp_entity"], scene["shp_type"], scene["shp_geom"], scene["shp_count"],
                              scene["shp_cat"], scene["shp_mask"])
