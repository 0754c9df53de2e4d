"""GPU: the drop-in headers used from a plain C++ program with std::string / int64 / enum / uint32 EntityIds and a
Grid registry (tests/cpp/api_check.cpp, built by colli2de_b200/build.py)."""
import os
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_api_check_program():
    exe = os.path.join(REPO, "colli2de_b200", "api_check")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr
