"""CPU: the oracle (and, when built, the unmodified reference) against the known-answer vectors restated from
the reference's own tests — this is what PINS the oracle (prompt section 3 / SURVEY.md 8(c))."""
import pytest

from tests import golden_cases as G, harness as H

BACKENDS = [n for n in ("oracle", "reference") if H.available(n)]


@pytest.mark.parametrize("name", BACKENDS)
def test_collide_known_answers(name):
    assert G.run_collide_cases(name) == 50


@pytest.mark.parametrize("name", BACKENDS)
def test_sweep_known_answers(name):
    assert G.run_sweep_cases(name) == 11


@pytest.mark.parametrize("name", BACKENDS)
def test_ray_known_answers(name):
    assert G.run_ray_cases(name) == 12


@pytest.mark.parametrize("name", BACKENDS)
def test_registry_known_answers(name):
    assert G.run_registry_cases(name) == 8


@pytest.mark.parametrize("name", BACKENDS)
def test_registry_ray_known_answers(name):
    assert G.run_registry_ray_cases(name) == 3


def test_oracle_is_available():
    assert "oracle" in BACKENDS, "oracle/libc2d_oracle.so missing: run __graft_entry__.build()"
