"""Seeded random inputs for the free-function parity tests (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import numpy as np

from tests import harness as H


def random_shapes(rng: np.random.Generator, n: int, types=None) -> H.ShapeBatch:
    stype = rng.integers(0, 4, n).astype(np.uint8) if types is None else np.asarray(types, np.uint8)
    geom = np.zeros((n, 16), np.float32)
    count = np.zeros(n, np.uint8)
    c = rng.uniform(-0.5, 0.5, (n, 2)).astype(np.float32)
    r = rng.uniform(0.2, 1.2, n).astype(np.float32)
    ang = rng.uniform(0, 2 * np.pi, n)
    half = rng.uniform(0.3, 1.5, n)
    d = np.stack([np.cos(ang), np.sin(ang)], 1) * half[:, None]
    m = stype == H.CIRCLE
    geom[m, 0:2], geom[m, 2] = c[m], r[m]
    m = stype == H.CAPSULE
    geom[m, 0:2], geom[m, 2:4], geom[m, 4] = (c - d)[m], (c + d)[m], (0.5 * r)[m]
    m = stype == H.SEGMENT
    geom[m, 0:2], geom[m, 2:4] = (c - d)[m], (c + d)[m]
    m = stype == H.POLYGON
    nv = rng.integers(3, 9, n)
    for k in range(8):
        theta = ang + 2 * np.pi * k / nv
        # slightly irregular but convex: radius fixed per polygon, uneven angular steps would risk concavity
        vx = c[:, 0] + r * np.cos(theta)
        vy = c[:, 1] + r * np.sin(theta)
        use = m & (k < nv)
        geom[use, 2 * k], geom[use, 2 * k + 1] = vx[use].astype(np.float32), vy[use].astype(np.float32)
    count[m] = nv[m]
    return H.ShapeBatch(stype, geom, count)


def random_xf3(rng: np.random.Generator, n: int, spread: float, rotate=True) -> np.ndarray:
    xf3 = np.zeros((n, 3), np.float32)
    xf3[:, :2] = rng.uniform(-spread, spread, (n, 2))
    if rotate:
        xf3[:, 2] = rng.uniform(-np.pi, np.pi, n)
    return xf3


def all_type_pairs(n_each: int):
    ta, tb = np.meshgrid(np.arange(4), np.arange(4), indexing="ij")
    return np.repeat(ta.ravel(), n_each).astype(np.uint8), np.repeat(tb.ravel(), n_each).astype(np.uint8)
