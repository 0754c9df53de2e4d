"""Secondary measurements for the other BASELINE.json configurations (not the driver's bench line):
cfg 1b (10 k circles + 1 k polygons), cfg 2 (100 k + 10 k mixed), cfg 4 (10 k bullets vs 1 M static polygons),
cfg 5 (batched rays vs a 4 M-shape scene), each next to the reference's Release build on the same host.

    python scripts/bench_configs.py [--rays 200000] [--frames 10] > profiles/r01_configs.json
"""
import argparse
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from colli2de_b200 import scenes  # noqa: E402
from tests import harness as H  # noqa: E402


def frames(backend, ids, amp, n_frames, seed=5, warm=2):
    rng = scenes.LCG(seed)
    times, pairs, dev = [], 0, []
    for f in range(warm + n_frames):
        backend.move_translate(ids, scenes.jitter(rng, len(ids), amp))
        sec = C.c_double(0.0)
        n = int(backend.lib.rn_get_colliding_pairs(backend.ctx, C.byref(sec)))
        backend._check()
        if f >= warm:
            times.append(sec.value)
            pairs += n
            dev.append(backend.stats()["ms_device"])
    return dict(ms_per_frame=1e3 * float(np.mean(times)), pairs_per_frame=pairs / n_frames,
                ms_device=float(np.mean(dev)) if any(dev) else None)


def pair_config(name, scene, ids, amp, n_frames, ref_frames):
    out = {"config": name, "shapes": int(len(scene["shp_entity"]))}
    b = H.Backend("b200")
    scenes.populate(b, scene)
    out["b200"] = frames(b, ids, amp, n_frames)
    if H.available("reference"):
        r = H.Backend("reference")
        scenes.populate(r, scene)
        out["reference"] = frames(r, ids, amp, ref_frames, warm=1)
        out["speedup_e2e"] = out["reference"]["ms_per_frame"] / out["b200"]["ms_per_frame"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=200_000)
    ap.add_argument("--frames", type=int, default=10)
    args = ap.parse_args()
    results = []
    sc = scenes.baseline_cfg1()
    results.append(pair_config("cfg1b: 10k dynamic circles + 1k static polygons", sc, np.arange(10_000, dtype=np.uint32), 0.5,
                               args.frames, 5))
    sc = scenes.mixed_scene(100_000, 10_000, 1600.0)
    results.append(pair_config("cfg2: 100k mixed dynamic + 10k static polygons", sc, np.arange(100_000, dtype=np.uint32), 0.5,
                               args.frames, 3))
    sc = scenes.bullet_scene(1_000_000, 10_000, 3200.0)
    results.append(pair_config("cfg4: 10k bullets (step <= 20) vs 1M static polygons", sc,
                               np.arange(1_000_000, 1_010_000, dtype=np.uint32), 20.0, args.frames, 2))
    # cfg 5: rays
    sc = scenes.mixed_scene(3_640_000, 360_000, 9535.0, seed=77)
    ray4, inf = scenes.rays(args.rays, 9535.0, seed=13)
    b = H.Backend("b200")
    scenes.populate(b, sc)
    b.raycast(ray4, inf)  # first call sizes the device / pinned buffers (GBs of cudaMalloc / cudaHostAlloc)
    t = time.perf_counter()
    off, hits, sec = b.raycast(ray4, inf, with_time=True)
    st = b.stats()
    entry = {"config": f"cfg5: {args.rays} rays (half finite len 50, half infinite) vs 4M shapes", "shapes": 4_000_000,
             "b200": {"rays_per_s_call": args.rays / sec, "rays_per_s_device": args.rays / (st["ms_device"] * 1e-3),
                      "hits": int(len(hits)), "ms_device": st["ms_device"], "ms_call": 1e3 * sec}}
    if H.available("reference"):
        r = H.Backend("reference")
        scenes.populate(r, sc)
        m = 1000
        _, rh, rsec = r.raycast(ray4[:m], inf[:m], with_time=True)
        entry["reference"] = {"rays_per_s": m / rsec, "sample": f"{m} rays looped through rayCast(ray)", "hits": int(len(rh))}
        entry["speedup_call"] = entry["b200"]["rays_per_s_call"] / entry["reference"]["rays_per_s"]
        entry["speedup_device"] = entry["b200"]["rays_per_s_device"] / entry["reference"]["rays_per_s"]
    results.append(entry)
    print(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
