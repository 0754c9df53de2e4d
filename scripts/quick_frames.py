"""Development helper: per-stage CUDA-event timings of getCollidingPairs on a BASELINE scene (B200 only)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from colli2de_b200 import scenes  # noqa: E402
from tests import harness as H  # noqa: E402

n_dyn = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_sta = n_dyn // 10
box = float(sys.argv[2]) if len(sys.argv) > 2 else 5000.0 * (n_dyn / 1e6) ** 0.5
clustered = len(sys.argv) > 3 and sys.argv[3] == "clustered"
sc = scenes.mixed_scene(n_dyn, n_sta, box, clustered=clustered)
b = H.Backend("b200")
t = time.time()
scenes.populate(b, sc)
print("populate s", time.time() - t, flush=True)
rng = scenes.LCG(1)
dyn = np.arange(n_dyn, dtype=np.uint32)
for f in range(6):
    d = scenes.jitter(rng, n_dyn, 0.5)
    t = time.time()
    b.move_translate(dyn, d)
    tm = time.time() - t
    t = time.time()
    p, sec = b.get_colliding_pairs(with_time=True)
    tq = time.time() - t
    print(f"frame {f}: pairs {len(p)} move_host {tm*1e3:.1f} ms  getCollidingPairs {sec*1e3:.2f} ms (py {tq*1e3:.1f})", flush=True)
