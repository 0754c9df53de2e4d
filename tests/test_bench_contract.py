"""CPU: the bench.py contract that can be checked without a GPU - the reference arm's JSON line (run on the small cfg 2
scene), the shared workload label of the two arms, and that the own arm FAILS LOUDLY without a CUDA device (there is no CPU
fallback behind the drop-in)."""
import json
import os
import subprocess
import sys

import pytest

from tests import harness as H

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=300):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], cwd=REPO, capture_output=True, text=True,
                          timeout=timeout, env=env)


@pytest.mark.skipif(not (H.available("reference") or H.available("oracle")), reason="needs a CPU checker library")
def test_reference_arm_line():
    r = _run("--impl", "reference", "--workload", "cfg2", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout  # stdout carries exactly one JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["pairs_per_frame"] > 1000
    sys.path.insert(0, REPO)
    import bench

    assert d["config"]["workload"] == bench.workload_label("cfg2", 100_000, 10_000)  # the own arm prints the same string


def test_own_arm_needs_cuda():
    r = _run("--workload", "cfg2", "--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--parity-frames", "0")
    assert r.returncode != 0
    assert "CUDA" in r.stderr and "no CPU fallback" in r.stderr
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]  # and no result line
