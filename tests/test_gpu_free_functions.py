"""GPU parity of the narrowphase free functions through the C ABI (c2d_gpu_*_batch) against the reference
(oracle/_ref when present, else the C oracle): collide / areColliding for all 16 type pairs, sweeps, shape
raycasts, computeAABB.  Bit-exact is expected and reported; the asserted bar is the north-star's:
booleans / pointCount exact, floats within 1e-5 relative."""
import numpy as np
import pytest

from tests import gen, harness as H, parity as P

pytestmark = pytest.mark.gpu

ORACLE = "reference" if H.available("reference") else "oracle"
SUBJECT = "b200"


def _assert_manifolds(exp, got, what):
    assert np.array_equal(exp["pointCount"], got["pointCount"]), f"{what}: pointCount differs"
    close = P.manifold_close(exp, got)
    assert close.all(), f"{what}: {int((~close).sum())} manifolds beyond 1e-5"
    return int((~P.manifold_bits_equal(exp, got)).sum())


def test_collide_all_type_pairs():
    rng = np.random.default_rng(1)
    ta, tb = gen.all_type_pairs(4000)
    n = len(ta)
    A, B = gen.random_shapes(rng, n, ta), gen.random_shapes(rng, n, tb)
    xa, xb = H.xf5_from_xf3(gen.random_xf3(rng, n, 1.2)), H.xf5_from_xf3(gen.random_xf3(rng, n, 1.2))
    em, ef = H.collide(ORACLE, A, xa, B, xb)
    gm, gf = H.collide(SUBJECT, A, xa, B, xb)
    assert np.array_equal(ef, gf)
    assert 0.15 < ef.mean() < 0.9
    nb = _assert_manifolds(em, gm, "collide")
    assert nb == 0, f"{nb} manifolds not bit-identical"


def test_collide_drifted_rotation():
    """Stored sin/cos that are NOT sinf/cosf(angle) (Rotation::operator+= quirk)."""
    rng = np.random.default_rng(2)
    ta, tb = gen.all_type_pairs(500)
    n = len(ta)
    A, B = gen.random_shapes(rng, n, ta), gen.random_shapes(rng, n, tb)
    xa, xb = H.xf5_from_xf3(gen.random_xf3(rng, n, 1.0)), H.xf5_from_xf3(gen.random_xf3(rng, n, 1.0))
    xa[:, 4] *= np.float32(0.9)
    xb[:, 3] *= np.float32(1.1)
    em, ef = H.collide(ORACLE, A, xa, B, xb)
    gm, gf = H.collide(SUBJECT, A, xa, B, xb)
    assert np.array_equal(ef, gf)
    assert _assert_manifolds(em, gm, "collide drifted") == 0


@pytest.mark.parametrize("mode", [1, 2])
def test_sweep_translation_only(mode):
    rng = np.random.default_rng(3 + mode)
    ta, tb = gen.all_type_pairs(600)
    n = len(ta)
    A, B = gen.random_shapes(rng, n, ta), gen.random_shapes(rng, n, tb)
    a0 = gen.random_xf3(rng, n, 4.0)
    a1 = a0.copy()
    a1[:, :2] += rng.uniform(-6, 6, (n, 2)).astype(np.float32)
    b0 = gen.random_xf3(rng, n, 2.0)
    b1 = b0.copy()
    b1[:, :2] += rng.uniform(-3, 3, (n, 2)).astype(np.float32)
    xa0, xa1, xb0, xb1 = (H.xf5_from_xf3(v) for v in (a0, a1, b0, b1))
    ef, em, eh = H.sweep(ORACLE, mode, A, xa0, xa1, B, xb0, xb1)
    gf, gm, gh = H.sweep(SUBJECT, mode, A, xa0, xa1, B, xb0, xb1)
    assert np.array_equal(eh, gh)
    assert 0.05 < eh.mean() < 0.95
    assert np.array_equal(ef[eh > 0].view(np.uint32), gf[gh > 0].view(np.uint32)), "sweep fractions differ"
    assert _assert_manifolds(em, gm, "sweep") == 0


def test_sweep_rotating_within_tolerance():
    """Rotating bullets: device sin/cos come from double precision, glibc's sinf is < 1 ulp — booleans may
    flip only on knife-edge samples; require >= 99.9% identical hit flags and close manifolds on agreement."""
    rng = np.random.default_rng(9)
    ta, tb = gen.all_type_pairs(300)
    n = len(ta)
    A, B = gen.random_shapes(rng, n, ta), gen.random_shapes(rng, n, tb)
    a0, b0 = gen.random_xf3(rng, n, 3.0), gen.random_xf3(rng, n, 1.5)
    a1 = a0 + np.concatenate([rng.uniform(-5, 5, (n, 2)), rng.uniform(-1.5, 1.5, (n, 1))], 1).astype(np.float32)
    xa0, xa1, xb0 = (H.xf5_from_xf3(v) for v in (a0, a1, b0))
    ef, em, eh = H.sweep(ORACLE, 1, A, xa0, xa1, B, xb0)
    gf, gm, gh = H.sweep(SUBJECT, 1, A, xa0, xa1, B, xb0)
    agree = eh == gh
    assert agree.mean() >= 0.999
    both = agree & (eh > 0) & (np.abs(ef - gf) < 1e-6)
    assert both.sum() > 0.9 * (eh > 0).sum()
    assert P.manifold_close(em[both], gm[both], rel=1e-4).mean() > 0.995


def test_raycast_shapes_and_ray_sweep():
    rng = np.random.default_rng(11)
    n = 40000
    S = gen.random_shapes(rng, n)
    xf = H.xf5_from_xf3(gen.random_xf3(rng, n, 1.0))
    ray = np.zeros((n, 4), np.float32)
    ray[:, :2] = rng.uniform(-4, 4, (n, 2))
    inf = (np.arange(n) % 2).astype(np.uint8)
    target = rng.uniform(-1.5, 1.5, (n, 2))
    d = target - ray[:, :2]
    ray[:, 2:] = np.where(inf[:, None] == 1, d / np.linalg.norm(d, axis=1, keepdims=True), target + d * 0.5)
    # a few degenerate rays
    ray[:50, 2:] = np.where(inf[:50, None] == 1, 0.0, ray[:50, :2])
    et, eh = H.raycast_shape(ORACLE, S, xf, ray, inf)
    gt, gh = H.raycast_shape(SUBJECT, S, xf, ray, inf)
    assert np.array_equal(eh, gh)
    assert 0.2 < eh.mean() < 0.95
    assert np.array_equal(et.view(np.uint32), gt.view(np.uint32)), "ray times not bit-identical"
    # bullet ray sweep, translation only
    xf1 = xf.copy()
    xf1[:, :2] += rng.uniform(-3, 3, (n, 2)).astype(np.float32)
    et, eh = H.sweep_ray(ORACLE, S, xf, xf1, ray, inf)
    gt, gh = H.sweep_ray(SUBJECT, S, xf, xf1, ray, inf)
    assert np.array_equal(eh, gh)
    assert np.array_equal(et.view(np.uint32), gt.view(np.uint32))


def test_compute_aabb():
    rng = np.random.default_rng(12)
    n = 50000
    S = gen.random_shapes(rng, n)
    xf = H.xf5_from_xf3(gen.random_xf3(rng, n, 100.0))
    e = H.compute_aabb(ORACLE, S, xf)
    g = H.compute_aabb(SUBJECT, S, xf)
    assert np.array_equal(e.view(np.uint32), g.view(np.uint32))
