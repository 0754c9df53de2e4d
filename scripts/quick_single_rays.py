"""Development helper: latency of the one-ray-per-call Registry::rayCast vs the reference (B200 box)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from colli2de_b200 import scenes  # noqa: E402
from tests import harness as H  # noqa: E402

os.environ["RN_RAYCAST_SINGLE"] = "1"
for n_shapes in (11_000, 110_000, 1_100_000):
    box = 5000.0 * (n_shapes / 1.1e6) ** 0.5
    sc = scenes.mixed_scene(int(n_shapes / 1.1), int(n_shapes / 11), box, seed=77)
    ray4, inf = scenes.rays(200, box, seed=13)
    for name in ("b200", "reference"):
        if not H.available(name):
            continue
        b = H.Backend(name)
        scenes.populate(b, sc)
        b.raycast(ray4[:10], inf[:10])
        for which, sel in (("finite", inf == 0), ("infinite", inf == 1)):
            t = time.perf_counter()
            off, hits = b.raycast(ray4[sel], inf[sel])
            dt = time.perf_counter() - t
            print(f"{n_shapes:8d} shapes {name:9s} {which:8s}: {1e6*dt/sel.sum():8.1f} us/ray, {len(hits)/sel.sum():6.1f} hits/ray", flush=True)
