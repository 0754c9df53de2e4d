"""Development helper: batched rayCast throughput for LIGHT rays (short finite rays, few hits each)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from colli2de_b200 import scenes  # noqa: E402
from tests import harness as H  # noqa: E402

n_shapes = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
length = float(sys.argv[3]) if len(sys.argv) > 3 else 5.0
box = 9535.0 * (n_shapes / 4e6) ** 0.5
sc = scenes.mixed_scene(int(n_shapes * 0.91), int(n_shapes * 0.09), box, seed=77)
b = H.Backend("b200")
scenes.populate(b, sc)
ray4, inf = scenes.rays(n_rays, box, seed=13, finite_len=length)
inf[:] = 0
d = ray4[:, 2:] - ray4[:, :2]
nrm = np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6)
ray4[:, 2:] = ray4[:, :2] + d / nrm * length
for rep in range(3):
    off, hits, sec = b.raycast(ray4, inf, with_time=True)
    st = b.stats()
    print(f"rays {n_rays} len {length} hits {len(hits)} leaf-box hits {int(st['candidates'])} call {sec*1e3:.1f} ms device "
          f"{st['ms_device']:.1f} ms ({n_rays/sec/1e6:.2f} Mrays/s) launches {int(st['kernel_launches'])}", flush=True)
