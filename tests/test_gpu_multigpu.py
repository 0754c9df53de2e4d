"""GPU, >= 2 devices: the multi-GPU group on real ranks over NCCL (tests/multigpu_check.py under torchrun).  Skipped on a
one-GPU box; the same decomposition with emulated ranks runs in tests/test_gpu_sharding.py."""
import os
import signal
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4])
def test_group_on_real_ranks(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(REPO, "tests", "multigpu_check.py")]
    # own session: on a timeout the whole process group goes (a killed torchrun would leave its ranks spinning on the GPUs)
    proc = subprocess.Popen(cmd, cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=300)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        raise AssertionError("multigpu_check.py timed out\n" + out[-2000:] + err[-3000:])
    assert proc.returncode == 0, out[-3000:] + err[-3000:]
    assert "MULTIGPU_OK" in out
