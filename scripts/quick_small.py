"""Development helper: per-call latency of the query entry points on a small scene."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from colli2de_b200 import scenes  # noqa: E402
from tests import harness as H  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
b = H.Backend("b200", 1, 120)
sc = scenes.repo_perf_scene(n)
scenes.populate(b, sc)
b.move_translate(sc["ent_id"], sc["first_move"])
sec = C.c_double()
for name, call in (("getCollidingPairs", lambda: b.lib.rn_get_colliding_pairs(b.ctx, C.byref(sec))),
                   ("countOnDevice", lambda: b.lib.rn_count_colliding_pairs_device(b.ctx, C.byref(sec))),
                   ("view", lambda: b.lib.rn_get_colliding_pairs_view(b.ctx, C.byref(sec), C.byref(C.c_void_p())))):
    for _ in range(20):
        call()
    ts = []
    for _ in range(200):
        call()
        ts.append(sec.value)
    st = b.stats()
    print(f"{name}: median {1e6*np.median(ts):.1f} us  min {1e6*min(ts):.1f} us  pairs {int(st['pairs'])} launches {int(st['kernel_launches'])} "
          f"ms_device {st['ms_device']*1e3:.1f} us", flush=True)
from colli2de_b200 import capi  # noqa: E402
ctx = b.native_context()
capi.set_profiling(ctx, True)
for _ in range(50):
    b.lib.rn_count_colliding_pairs_device(b.ctx, C.byref(sec))
for k, (cnt, ms) in sorted(capi.kernel_profile(ctx).items(), key=lambda kv: -kv[1][1]):
    print(f"  {k}: {cnt} launches, {1e3*ms/cnt:.1f} us each")
