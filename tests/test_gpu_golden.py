"""GPU: the reference's known-answer vectors (tests/golden_cases.py) through the C ABI / the drop-in Registry."""
import pytest

from tests import golden_cases as G

pytestmark = pytest.mark.gpu


def test_collide_known_answers():
    assert G.run_collide_cases("b200") == 50


def test_sweep_known_answers():
    assert G.run_sweep_cases("b200") == 11


def test_ray_known_answers():
    assert G.run_ray_cases("b200") == 12


def test_registry_known_answers():
    assert G.run_registry_cases("b200") == 8


def test_registry_ray_known_answers():
    assert G.run_registry_ray_cases("b200") == 3
