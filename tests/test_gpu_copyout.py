"""The copy-out entry points of the C ABI (c2d_gpu_collide_begin / prefault / collide_end / collide_fetch,
c2d_gpu_raycast_begin / raycast_fetch): records delivered straight into caller memory by the library's helper threads
must equal the ones c2d_gpu_collide / c2d_gpu_raycast leave in the context's pinned buffer, for results below and above
the 4 MB threshold of the threaded path, with and without helper threads; misuse fails loudly."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from colli2de_b200 import capi, scenes
from tests import harness as H

pytestmark = pytest.mark.gpu

STATS = C.c_byte * 256


def _lib():
    lib = capi.lib()
    lib.c2d_gpu_collide.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_void_p]
    lib.c2d_gpu_collide_begin.argtypes = [C.c_void_p]
    lib.c2d_gpu_prefault.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.c2d_gpu_collide_end.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]
    lib.c2d_gpu_collide_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.c2d_gpu_raycast.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]
    lib.c2d_gpu_raycast_begin.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.POINTER(C.c_uint64),
                                          C.POINTER(C.c_void_p), C.c_void_p]
    lib.c2d_gpu_raycast_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.c2d_gpu_last_error.restype = C.c_char_p
    lib.c2d_gpu_last_error.argtypes = [C.c_void_p]
    return lib


def _sorted(recs):
    out = recs.copy()
    out["_pad"] = 0
    return out[np.lexsort((out["shapeB"], out["shapeA"], out["entityB"], out["entityA"]))]


def _pinned_pairs(lib, ctx):
    ptr, n = C.c_void_p(), C.c_uint64()
    assert lib.c2d_gpu_collide(ctx, C.byref(ptr), C.byref(n), None) == 0, lib.c2d_gpu_last_error(ctx)
    return np.frombuffer(C.string_at(ptr, n.value * 84), dtype=H.PAIR_DTYPE).copy()


def _fetched_pairs(lib, ctx, prefault=True):
    assert lib.c2d_gpu_collide_begin(ctx) == 0, lib.c2d_gpu_last_error(ctx)
    guess = np.empty(1 << 23, np.uint8)  # caller storage reserved while the GPU works
    if prefault:
        assert lib.c2d_gpu_prefault(ctx, guess.ctypes.data_as(C.c_void_p), guess.nbytes) == 0
    n = C.c_uint64()
    assert lib.c2d_gpu_collide_end(ctx, C.byref(n), None) == 0, lib.c2d_gpu_last_error(ctx)
    out = np.zeros(n.value, H.PAIR_DTYPE)
    assert lib.c2d_gpu_collide_fetch(ctx, out.ctypes.data_as(C.c_void_p), n.value, None) == 0, lib.c2d_gpu_last_error(ctx)
    return out


@pytest.mark.parametrize("n_dyn,box", [(3000, 160.0), (260_000, 900.0)])
def test_collide_fetch_equals_pinned_records(n_dyn, box):
    lib = _lib()
    b = H.Backend("b200")
    sc = scenes.mixed_scene(n_dyn, n_dyn // 10, box, seed=12)
    scenes.populate(b, sc)
    ctx = b.native_context()
    want = _pinned_pairs(lib, ctx)
    got = _fetched_pairs(lib, ctx)
    assert len(want) == len(got) and (len(got) * 84 > (4 << 20)) == (n_dyn > 100_000)
    assert _sorted(want).tobytes() == _sorted(got).tobytes()
    # a bracket abandoned after begin is drained by the next begin; results stay right
    assert lib.c2d_gpu_collide_begin(ctx) == 0
    again = _fetched_pairs(lib, ctx, prefault=False)
    assert _sorted(again).tobytes() == _sorted(want).tobytes()
    # misuse: more records than the frame produced, end without begin
    n = C.c_uint64()
    assert lib.c2d_gpu_collide_fetch(ctx, again.ctypes.data_as(C.c_void_p), len(again) + 1, None) != 0
    assert b"more records" in lib.c2d_gpu_last_error(ctx)
    assert lib.c2d_gpu_collide_end(ctx, C.byref(n), None) != 0
    assert b"without c2d_gpu_collide_begin" in lib.c2d_gpu_last_error(ctx)


def test_raycast_fetch_equals_pinned_records():
    lib = _lib()
    b = H.Backend("b200")
    sc = scenes.mixed_scene(60_000, 6_000, 1200.0, seed=4)
    scenes.populate(b, sc)
    ctx = b.native_context()
    for n_rays in (40, 30_000):  # latency path (results already on the host) and throughput path (results in HBM)
        ray4, inf = scenes.rays(n_rays, 1200.0, seed=6)
        rays = np.zeros(n_rays, dtype=[("r", "<f4", (4,)), ("mask", "<u8"), ("inf", "<u4"), ("pad", "<u4")])
        rays["r"], rays["mask"], rays["inf"] = ray4, ~np.uint64(0), inf
        hits, offs = C.c_void_p(), C.c_void_p()
        assert lib.c2d_gpu_raycast(ctx, n_rays, rays.ctypes.data_as(C.c_void_p), C.byref(hits), C.byref(offs), None) == 0
        want_off = np.frombuffer(C.string_at(offs, 8 * (n_rays + 1)), np.uint64).copy()
        want = np.frombuffer(C.string_at(hits, int(want_off[-1]) * 32), H.HIT_DTYPE).copy()
        total = C.c_uint64()
        assert lib.c2d_gpu_raycast_begin(ctx, n_rays, rays.ctypes.data_as(C.c_void_p), 0, C.byref(total), C.byref(offs), None) == 0
        got_off = np.frombuffer(C.string_at(offs, 8 * (n_rays + 1)), np.uint64).copy()
        got = np.zeros(total.value, H.HIT_DTYPE)
        assert lib.c2d_gpu_raycast_fetch(ctx, got.ctypes.data_as(C.c_void_p), total.value, None) == 0, lib.c2d_gpu_last_error(ctx)
        assert total.value == want_off[-1] > n_rays and np.array_equal(want_off, got_off)
        assert want.tobytes() == got.tobytes()
        assert lib.c2d_gpu_raycast_fetch(ctx, got.ctypes.data_as(C.c_void_p), total.value + 1, None) != 0


def test_copy_threads_can_be_turned_off():
    """C2D_GPU_COPY_THREADS=1: no helper threads, the calling thread does the whole copy (read at first use, so a fresh
    process is needed)."""
    code = (
        "import numpy as np\n"
        "from colli2de_b200 import scenes\n"
        "from tests import harness as H\n"
        "b = H.Backend('b200'); sc = scenes.mixed_scene(260000, 26000, 900.0, seed=12); scenes.populate(b, sc)\n"
        "p = b.get_colliding_pairs(); v, _ = b.get_colliding_pairs_view()\n"
        "assert len(p) * 84 > (4 << 20) and len(p) == len(v)\n"
        "key = lambda r: np.sort(r['shapeA'].astype(np.uint64) << np.uint64(32) | r['shapeB'])\n"
        "assert np.array_equal(key(p), key(v)); print('ok', len(p))\n")
    env = dict(os.environ, C2D_GPU_COPY_THREADS="1", PYTHONPATH=H.REPO)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=H.REPO, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
