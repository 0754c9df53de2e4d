"""GPU: PartitioningMethod::Grid.  The reference's grid-of-BVHs reports duplicates; the drop-in
Registry<Id, Grid> must return the same UNIQUE unordered pair set (canonical orientation) with matching manifolds,
and rayCast must reproduce the grid's 50-cell truncation of infinite rays (BroadPhaseTree.hpp:22,531-537)."""
import numpy as np
import pytest

from colli2de_b200 import scenes
from tests import harness as H, parity as P

pytestmark = pytest.mark.gpu

SUBJECT = "b200"


@pytest.mark.skipif(not H.available("reference"), reason="the C oracle does not restate the grid; needs oracle/_ref")
def test_grid_registry_pairs_and_rays():
    cell = 40
    sc = scenes.mixed_scene(12000, 1200, 520.0, seed=21)
    sc["ent_body"][::45] = 2
    ref, gpu = H.Backend("reference", 1, cell), H.Backend(SUBJECT, 1, cell)
    scenes.populate(ref, sc)
    scenes.populate(gpu, sc)
    idx = P.SceneIndex(sc)
    rng = scenes.LCG(4)
    mov = np.nonzero(sc["ent_body"] != 0)[0].astype(np.uint32)
    for frame in range(3):
        d = scenes.jitter(rng, len(mov), 0.8)
        ref.move_translate(mov, d)
        gpu.move_translate(mov, d)
        rp, gp = ref.get_colliding_pairs(), gpu.get_colliding_pairs()
        res = P.compare_pairs("reference", ref, idx, rp, gp, max_ambiguous=8)
        assert res["expected"] > 2000 and res["not_bit_identical"] == 0, res
        assert len(rp) >= res["expected"]  # the reference's list has the duplicates
    # rays: finite rays as with None; infinite rays only reach 50 cells = 2000 units... use a small cell so the
    # truncation actually cuts: cell 40 -> 2000 > scene; so also test a registry with cell 4 (reach 200)
    for c in (cell, 4):
        ref, gpu = H.Backend("reference", 1, c), H.Backend(SUBJECT, 1, c)
        small = scenes.mixed_scene(3000, 300, 520.0, seed=22)
        scenes.populate(ref, small)
        scenes.populate(gpu, small)
        ray4, inf = scenes.rays(600, 520.0, seed=9)
        eo, eh = ref.raycast(ray4, inf)
        go, gh = gpu.raycast(ray4, inf)
        bad = 0
        for i in range(600):
            a, b = eh[int(eo[i]):int(eo[i + 1])], gh[int(go[i]):int(go[i + 1])]
            bad += not (len(a) == len(b) and a.tobytes() == b.tobytes())
        assert bad <= 6, f"cell {c}: {bad} of 600 rays differ"
        assert len(eh) > 600
