"""CPU: the C oracle (oracle/c2d_oracle.c) against the UNMODIFIED reference compiled into oracle/_ref, on the
same seeded inputs the GPU parity tests use — random shape pairs of all 16 type pairs, sweeps, shape raycasts,
computeAABB, and scripted registry scenes (translate / transform moves, teleports, bullets, activation,
removal).  Bit-exact (both run the host's libm).  Skipped where /root/reference was never available to build
oracle/_ref (the oracle is then pinned by tests/test_oracle_golden.py alone)."""
import pytest

from tests import harness as H
from tests import test_gpu_free_functions as F
from tests import test_gpu_registry_pairs as R

pytestmark = pytest.mark.skipif(not (H.available("reference") and H.available("oracle")),
                                reason="needs oracle/_ref and oracle/libc2d_oracle.so")


@pytest.fixture(autouse=True)
def oracle_as_subject(monkeypatch):
    for mod in (F, R):
        monkeypatch.setattr(mod, "ORACLE", "reference")
        monkeypatch.setattr(mod, "SUBJECT", "oracle")


def test_collide_all_type_pairs():
    F.test_collide_all_type_pairs()


def test_collide_drifted_rotation():
    F.test_collide_drifted_rotation()


@pytest.mark.parametrize("mode", [1, 2])
def test_sweeps(mode):
    F.test_sweep_translation_only(mode)
    F.test_sweep_rotating_within_tolerance()


def test_raycasts_and_aabb():
    F.test_raycast_shapes_and_ray_sweep()
    F.test_compute_aabb()


def test_registry_scenes():
    R.test_baseline_cfg1_static_frames()
    R.test_mixed_with_bullets_translate_frames()
    R.test_transform_moves_teleports_removal_activation()
    R.test_category_filter_and_same_entity()


def test_secondary_oracles(monkeypatch):
    """tight boxes, leaf boxes (fat-leaf rule) and the candidate set of the C oracle against the reference's private
    state, bit for bit over the scripted scene of tests/test_gpu_secondary_oracles.py."""
    from tests import test_gpu_secondary_oracles as S

    monkeypatch.setattr(S, "SUBJECT", "oracle")
    totals = S.run_scripted_scene(n_dyn=3000, n_static=800, side=160.0, frames=5)
    assert all(c > 4000 for _, c in totals), totals
    S.test_boxes_and_candidates_cfg1b_frames()
