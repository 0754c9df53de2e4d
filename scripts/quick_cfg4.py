"""Development helper: per-kernel times of the bullet configuration (cfg 4)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from colli2de_b200 import capi, scenes  # noqa: E402
from tests import harness as H  # noqa: E402

sc = scenes.bullet_scene(1_000_000, 10_000, 3200.0)
b = H.Backend("b200")
scenes.populate(b, sc)
ctx = b.native_context()
bul = np.arange(1_000_000, 1_010_000, dtype=np.uint32)
rng = scenes.LCG(5)
for f in range(3):
    b.move_translate(bul, scenes.jitter(rng, len(bul), 20.0))
    b.get_colliding_pairs()
capi.set_profiling(ctx, True)
capi.kernel_profile(ctx, reset=True)
for f in range(5):
    b.move_translate(bul, scenes.jitter(rng, len(bul), 20.0))
    p = b.get_colliding_pairs()
    st = b.stats()
print({k: round(v, 3) for k, v in st.items()})
for k, (c, ms) in sorted(capi.kernel_profile(ctx).items(), key=lambda kv: -kv[1][1]):
    print(f"{k:45s} {c:3d} {ms / c:8.4f} ms")
