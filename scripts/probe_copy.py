"""Probe (GPU box): where does the result download spend its time?  One cfg 3 frame, then c2d_gpu_collide_fetch of the
records alone, repeated, next to a plain pinned cudaMemcpy of the same bytes."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from colli2de_b200 import capi, scenes  # noqa: E402
from tests import harness as H  # noqa: E402

sc = scenes.mixed_scene(1_000_000, 100_000, 5000.0, clustered=True)
b = H.Backend("b200")
scenes.populate(b, sc)
ctx = b.native_context()
n, _ = b.count_pairs_device()
print("pairs", n, "bytes", n * 84)
lib = capi.lib()
lib.c2d_gpu_collide_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
for trial in range(3):
    dst = np.empty(n * 84, np.uint8)
    dst[::4096] = 0  # faulted in
    for rep in range(4):
        t = time.perf_counter()
        rc = lib.c2d_gpu_collide_fetch(ctx, dst.ctypes.data, n, None)
        dt = time.perf_counter() - t
        print(f"fetch into faulted pageable: {dt * 1e3:.3f} ms  {n * 84 / dt / 1e9:.1f} GB/s rc={rc}")
src = torch.empty(n * 84, dtype=torch.uint8, device="cuda")
pin = torch.empty(n * 84, dtype=torch.uint8).pin_memory()
for rep in range(4):
    torch.cuda.synchronize()
    t = time.perf_counter()
    pin.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"plain D2H into pinned: {dt * 1e3:.3f} ms  {n * 84 / dt / 1e9:.1f} GB/s")
host = torch.empty(n * 84, dtype=torch.uint8)
for rep in range(3):
    t = time.perf_counter()
    host.copy_(pin)
    dt = time.perf_counter() - t
    print(f"single-thread memcpy pinned -> pageable: {dt * 1e3:.3f} ms  {n * 84 / dt / 1e9:.1f} GB/s")
print("cpus", os.cpu_count())
