"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck): pairs (LBVH and grid), bullets,
rays (batch, first-hit, single), per-entity queries, staged moves, removal / recycling, overflow retries."""
import sys

import numpy as np

sys.path.insert(0, ".")
from colli2de_b200 import scenes  # noqa: E402
from tests import harness as H  # noqa: E402

for method in (0, 1):
    sc = scenes.mixed_scene(3000, 300, 120.0, seed=3, clustered=bool(method))
    sc["ent_body"][::25] = 2
    b = H.Backend("b200", method, 16)
    scenes.populate(b, sc)
    rng = scenes.LCG(1)
    mov = np.nonzero(sc["ent_body"] != 0)[0].astype(np.uint32)
    h = b.stage_translations(mov, scenes.jitter(rng, len(mov), 0.4))
    for f in range(3):
        b.apply_staged(h)
        n, _ = b.count_pairs_device()
        b.move_translate(mov, scenes.jitter(rng, len(mov), 0.5))
        p = b.get_colliding_pairs()
        print("method", method, "frame", f, "pairs", len(p), n, flush=True)
    b.release_staged(h)
    ray4, inf = scenes.rays(300, 120.0, seed=2)
    off, hits = b.raycast(ray4, inf)
    found, fh = b.first_hit_raycast(ray4[:100], inf[:100])
    col = sum(len(b.get_collisions(int(e))) for e in range(0, 3300, 97))
    ac = b.are_colliding(np.arange(0, 500, dtype=np.uint32), np.arange(1, 501, dtype=np.uint32))
    b.remove_entities(np.arange(5, 3300, 9, dtype=np.uint32))
    p = b.get_colliding_pairs()
    print("rays", len(hits), "first", int(found.sum()), "collisions", col, "areColliding", int(ac.sum()), "after removal", len(p), flush=True)
print("sanitize run done")
