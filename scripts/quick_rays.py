"""Development helper: batched rayCast throughput on the cfg-5 scene (B200 only)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from colli2de_b200 import scenes  # noqa: E402
from tests import harness as H  # noqa: E402

n_shapes = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
box = 9535.0 * (n_shapes / 4e6) ** 0.5
sc = scenes.mixed_scene(int(n_shapes * 0.91), int(n_shapes * 0.09), box, seed=77)
b = H.Backend("b200")
scenes.populate(b, sc)
ray4, inf = scenes.rays(n_rays, box, seed=13)
for rep in range(3):
    t = time.time()
    off, hits, sec = b.raycast(ray4, inf, with_time=True)
    st = b.stats()
    print(f"rays {n_rays} hits {len(hits)} leaf-box hits {int(st['candidates'])} call {sec*1e3:.1f} ms device {st['ms_device']:.1f} ms "
          f"({n_rays/sec/1e6:.2f} Mrays/s) launches {int(st['kernel_launches'])}", flush=True)
