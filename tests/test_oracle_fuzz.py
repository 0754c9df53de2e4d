"""CPU: the C oracle against the unmodified reference under the random API scripts of tests/test_gpu_fuzz.py (every
mutator of the public API, multi-shape entities, ShapeId recycling, both partitioning methods) — pins the oracle's
registry semantics beyond the scripted scenarios of test_oracle_vs_reference.py."""
import numpy as np
import pytest

from tests import harness as H
from tests import test_gpu_fuzz as F

needs = pytest.mark.skipif(not (H.available("reference") and H.available("oracle")), reason="needs oracle/_ref and the C oracle")


@needs
@pytest.mark.parametrize("seed", range(12))
def test_oracle_follows_reference_through_random_scripts(seed, monkeypatch):
    monkeypatch.setenv("C2D_DRYRUN_CANON", "1")  # the oracle reports same-tree pairs in its own orientation
    rng = np.random.default_rng(1000 + seed)
    method, cell = (1, 16) if seed % 3 == 2 else (0, 120)
    ref, sub = H.Backend("reference"), H.Backend("oracle", method, cell)
    ref_grid = H.Backend("reference", 1, cell) if method == 1 else None
    backends = (ref, sub) + ((ref_grid,) if ref_grid is not None else ())
    w = F.World()
    F.add_entities(rng, w, backends, int(rng.integers(150, 500)))
    total = 0
    for _ in range(7):
        F.random_step(rng, w, backends, teleports=method == 0)
        total += F.check_queries(rng, w, ref, sub, ref_grid)
        backends = tuple(b for b in backends if not getattr(b, "broken", False))  # a thrown reference Grid registry is out
    assert total > 100
