"""The committed golden fixtures (tests/golden/*.npz, produced from the UNMODIFIED reference by
tests/golden/make_golden.py) replayed on the C oracle (CPU) and on the CUDA path (GPU).  They keep the parity
pinned on a box where neither /root/reference nor a prebuilt oracle/_ref exists."""
import os

import numpy as np
import pytest

from colli2de_b200 import scenes
from tests import harness as H, parity as P
from tests.golden import make_golden as MG

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_functions(name):
    z = np.load(os.path.join(GOLD, "free_functions_v1.npz"))
    A = H.ShapeBatch(z["typeA"], z["geomA"], z["countA"])
    B = H.ShapeBatch(z["typeB"], z["geomB"], z["countB"])
    man, flag = H.collide(name, A, z["xa"], B, z["xb"])
    assert np.array_equal(flag, z["colliding"])
    assert P.manifold_bits_equal(man, z["manifold"]).all()
    for mode in (1, 2):
        f, m, h = H.sweep(name, mode, A, z["xa"], z["xa_end"], B, z["xb"], z["xb_end"])
        assert np.array_equal(h, z[f"sweep{mode}_hit"])
        hit = h > 0
        assert np.array_equal(f[hit].view(np.uint32), z[f"sweep{mode}_fraction"][hit].view(np.uint32))
        assert P.manifold_bits_equal(m[hit], z[f"sweep{mode}_manifold"][hit]).all()
    t, h = H.raycast_shape(name, A, z["xa"], z["ray"], z["infinite"])
    assert np.array_equal(h, z["ray_hit"]) and np.array_equal(t.view(np.uint32), z["ray_times"].view(np.uint32))
    t, h = H.sweep_ray(name, A, z["xa"], z["xa_end"], z["ray"], z["infinite"])
    assert np.array_equal(h, z["raysweep_hit"]) and np.array_equal(t.view(np.uint32), z["raysweep_times"].view(np.uint32))
    assert np.array_equal(H.compute_aabb(name, A, z["xa"]).view(np.uint32), z["aabb"].view(np.uint32))
    return len(flag)


def _registry_scene(name):
    z = np.load(os.path.join(GOLD, "registry_scene_v1.npz"))
    sc = MG.scripted_scene()
    b = H.Backend(name)
    scenes.populate(b, sc)
    rng = scenes.LCG(77)
    total = 0
    for frame in range(5):
        for method, args in MG.script(frame, sc, rng):
            getattr(b, method)(*args)
        got = P.sort_pairs(b.get_colliding_pairs())
        exp, amb = z[f"pairs_{frame}"], z[f"ambiguous_{frame}"]
        kg, ke, ka = P.pair_keys(got), P.pair_keys(exp), P.pair_keys(amb)
        assert len(np.setdiff1d(ke, kg)) == 0, f"frame {frame}: missing pairs"
        assert len(np.setdiff1d(np.setdiff1d(kg, ke), ka)) == 0, f"frame {frame}: unexpected pairs"
        common = got[np.isin(kg, ke)]
        assert P.manifold_bits_equal(common, exp).all(), f"frame {frame}: manifolds differ"
        off, hits = b.raycast(z["ray4"], z["ray_infinite"], z["ray_masks"])
        same_rays = np.array_equal(off, z[f"ray_offsets_{frame}"]) and hits.tobytes() == z[f"ray_hits_{frame}"].tobytes()
        assert same_rays, f"frame {frame}: ray hits differ"
        total += len(exp)
    return total


@pytest.mark.skipif(not H.available("oracle"), reason="needs oracle/libc2d_oracle.so")
def test_oracle_matches_golden_fixtures():
    assert _free_functions("oracle") == 16 * 120
    assert _registry_scene("oracle") > 4000


@pytest.mark.gpu
def test_cuda_path_matches_golden_fixtures():
    assert _free_functions("b200") == 16 * 120
    assert _registry_scene("b200") > 4000
