"""CPU: the C-ABI library loads and exports every symbol include/c2d_gpu.h declares (no compute calls here),
and the product has no CPU fallback: creating a context without a CUDA device fails loudly."""
import ctypes as C
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "c2d_gpu.h")
LIB = os.path.join(REPO, "colli2de_b200", "libc2d_gpu.so")
RUNNER = os.path.join(REPO, "colli2de_b200", "libc2d_b200_runner.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(c2d_gpu_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("c2d_gpu_create", "c2d_gpu_collide", "c2d_gpu_raycast", "c2d_gpu_shape_add",
                 "c2d_gpu_shape_translate", "c2d_gpu_shape_move", "c2d_gpu_collide_batch"):
        assert must in names
    assert len(names) >= 25


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "libc2d_gpu.so missing: run __graft_entry__.build()"
    lib = C.CDLL(LIB)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared in c2d_gpu.h but not exported: {missing}"


def test_capi_export_list_matches_header():
    from colli2de_b200 import capi

    assert sorted(capi.EXPORTS) == declared_symbols()


def test_runner_built_against_the_drop_in_headers_loads():
    assert os.path.exists(RUNNER)
    lib = C.CDLL(RUNNER)
    lib.rn_backend_name.restype = C.c_char_p
    assert lib.rn_backend_name() == b"b200"


def test_no_cpu_fallback_without_a_device():
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    lib = C.CDLL(LIB)
    lib.c2d_gpu_last_error.restype = C.c_char_p
    lib.c2d_gpu_last_error.argtypes = [C.c_void_p]
    ctx = C.c_void_p()
    rc = lib.c2d_gpu_create(-1, C.byref(ctx))
    assert rc != 0 and not ctx.value
    assert b"no CPU fallback" in lib.c2d_gpu_last_error(None)
    # and through the drop-in Registry: the constructor throws, surfaced as the runner's error string
    from tests import harness as H

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        H.Backend("b200")
