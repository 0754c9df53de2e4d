#!/usr/bin/env python
"""Benchmark-report compatibility.

Re-runs the Registry benchmarks of the reference's own suite (tests/registry/test_registry_performance.cpp:15-290 and
tests/registry/test_registry_raycast_performance.cpp:15-80: same scenes, same names, same thresholds) against any
backend of tests/harness.py and writes `benchmarks--<date>--<time>.md` in the table format of the reference's
exporter (tests/utils/Performance.hpp:154-214, file name from tests/main.cpp:39-43), so that the reference's
scripts/compare_tests.sh (compare_tests_parse.py / compare_tests_format.py) can diff a CPU run against a GPU run:

    python scripts/reference_benchmarks.py --backend reference --out bench_out/cpu
    python scripts/reference_benchmarks.py --backend b200      --out bench_out/gpu      # needs a B200
    <reference>/scripts/compare_tests.sh bench_out/cpu/benchmarks--*.md bench_out/gpu/benchmarks--*.md

Timing: the clock brackets the same Registry calls as the reference's BENCHMARK_FUNCTION (reset excluded); queries are
timed inside the native runner, insert/move batches around one ctypes call that loops the per-entity API natively.
Catch2 picks its own sample count and warms the function up first; here it is --samples (default 10) after --warmup
(default 1) untimed calls, each sample one call like in the reference.
The B200 Registry defers device work to the next query, so "Insert"/"Move" measure its host-side logging only, and
each query sample that follows a reset also pays the replay of that log.
"""
from __future__ import annotations

import argparse
import datetime
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from colli2de_b200.scenes import LCG  # noqa: E402
from tests import harness  # noqa: E402

NONE, GRID = 0, 1
STATIC, DYNAMIC, BULLET = 0, 1, 2


def random_circles(n, lo, hi, seed):
    """generateRandomCircles (tests/utils/Random.hpp:62-76): centres only, the radius is a constant."""
    return LCG(seed).next_vec2(lo, hi, n)


def random_translations(n, lo, hi, seed):
    """generateRandomTranslations: n (x, y) pairs from the same LCG."""
    return LCG(seed).next_vec2(lo, hi, n)


def circle_geom(n, radius, centres=None):
    g = np.zeros((n, 16), np.float32)
    if centres is not None:
        g[:, :2] = centres
    g[:, 2] = radius
    return g


def xf3(centres):
    x = np.zeros((len(centres), 3), np.float32)
    x[:, :2] = centres
    return x


class Report:
    def __init__(self, samples, name_filter, warmup=1):
        self.samples, self.filter, self.rows, self.warmup = samples, name_filter, [], warmup

    def run(self, name, threshold_ms, func, reset=None):
        if self.filter not in name:
            return
        times = []
        # Catch2 runs the benchmarked function for a warm-up period (benchmark-warmup-time, 100 ms by default) before it
        # samples; here: --warmup untimed calls (default 1)
        for _ in range(self.warmup):
            if reset:
                reset()
            func()
        for _ in range(self.samples):
            if reset:
                reset()
            times.append(1e3 * func())
        self.rows.append((name, threshold_ms, times))
        avg = sum(times) / len(times)
        print(f"{name}: avg {avg:.3f} ms (threshold {threshold_ms} ms)  samples " + " ".join(f"{t:.3f}" for t in times), file=sys.stderr)

    def export(self, folder):
        """exportBenchmarks (tests/utils/Performance.hpp:154-214): sorted by average, 3 decimals, padded columns."""
        if not self.rows:
            return None
        rows = sorted(self.rows, key=lambda r: sum(r[2]) / len(r[2]))
        width = max([len("Benchmark Name")] + [len(r[0]) for r in rows])
        max_ms = max(max(max(r[2]), r[1]) for r in rows)
        digits = len(str(int(max_ms))) + 4
        lines = [f"{'Benchmark Name':<{width}} | Status | Threshold (ms) | Avg (ms) | Min (ms) | Max (ms) |",
                 f"{'-' * width} | {'-' * 6} | {'-' * 14} | {'-' * 8} | {'-' * 8} | {'-' * 8} |"]
        for name, thr, t in rows:
            avg = sum(t) / len(t)
            cells = [f"{v:{digits}.3f}" for v in (thr, avg, min(t), max(t))]
            lines.append(f"{name:<{width}} | {'PASS' if avg <= thr else 'FAIL':<6} | {cells[0]:<14} | {cells[1]:<8} | "
                         f"{cells[2]:<8} | {cells[3]:<8} |")
        os.makedirs(folder, exist_ok=True)
        stamp = datetime.datetime.now().strftime("%Y-%m-%d--%H-%M-%S")
        path = os.path.join(folder, f"benchmarks--{stamp}.md")
        with open(path, "w") as f:
            f.write("\n".join(lines) + "\n")
        return path


def timed(call):
    t0 = time.perf_counter()
    call()
    return time.perf_counter() - t0


def bench_insert(rep, backend, seed):
    centres = random_circles(100_000, -50.0, 50.0, seed)
    for n, label, thr in ((10_000, "10k", 15.0), (100_000, "100k", 150.0)):
        ids = np.arange(n, dtype=np.uint32)

        def insert():
            b = harness.Backend(backend, GRID, 64)
            body = np.full(n, DYNAMIC, np.uint8)

            def work():
                b.create_entities(ids, body, xf3(centres[:n]))
                b.add_shapes(ids, np.zeros(n, np.uint8), circle_geom(n, 1.0, centres[:n]))
            dt = timed(work)
            b.close()
            return dt
        rep.run(f"Registry | Insert {label} entities and shapes", thr, insert)


def bench_move(rep, backend, seed):
    centres = random_circles(100_000, -200.0, 200.0, seed)
    deltas = random_translations(100_000, -4.0, 4.0, seed + 1)
    b = harness.Backend(backend, GRID, 120)
    for n, label, thr in ((10_000, "10k", 5.0), (100_000, "100k", 50.0)):
        ids = np.arange(n, dtype=np.uint32)
        d3 = xf3(deltas[:n])

        def reset():
            b.clear()
            b.create_entities(ids, np.full(n, DYNAMIC, np.uint8), xf3(centres[:n]))
            b.add_shapes(ids, np.zeros(n, np.uint8), circle_geom(n, 1.0, centres[:n]))
        reset()
        rep.run(f"Registry | Move {label} entities", thr, lambda: timed(lambda: b.move_transform(ids, d3)), reset)
    b.close()


def query_scene(b, n, centres, all_bullets):
    i = np.arange(n)
    if all_bullets:
        body = np.full(n, BULLET, np.uint8)
    else:
        bullet = (i % 1000) == 0
        body = np.where((~bullet) & ((i % 3) == 0), STATIC, np.where(bullet, BULLET, DYNAMIC)).astype(np.uint8)
    ids = i.astype(np.uint32)
    b.clear()
    b.create_entities(ids, body, xf3(centres[:n]))
    b.add_shapes(ids, np.zeros(n, np.uint8), circle_geom(n, 12.0))
    b.move_translate(ids, np.tile(np.float32([3.0, 3.0]), (n, 1)))


def bench_all_pairs(rep, backend, seed):
    centres = random_circles(100_000, -1920.0, 3840.0, seed)
    b = harness.Backend(backend, GRID, 120)
    for kind, all_bullets, thresholds in (("entities", False, (0.030, 3.0, 150.0)), ("bullets", True, (0.030, 5.0, 250.0))):
        for (n, label), thr in zip(((1_000, "1k"), (10_000, "10k"), (100_000, "100k")), thresholds):
            query_scene(b, n, centres, all_bullets)
            rep.run(f"Registry | Find all colliding pairs among {label} {kind}", thr,
                    lambda: b.get_colliding_pairs(with_time=True)[1])
    b.close()


def bench_single_queries(rep, backend, seed):
    centres = random_circles(100_000, -1920.0, 3840.0, seed)
    b = harness.Backend(backend, GRID, 120)
    first = np.arange(1000, dtype=np.uint32)
    names = ("Registry | Find collisions 1-by-1 among 1k entities",
             "Registry | Find collisions 1-by-1 entity among 10k entities",
             "Registry | Find collisions 1-by-1 entity among 100k entities")
    for n, name, thr in zip((1_000, 10_000, 100_000), names, (1.0, 2.0, 5.0)):
        query_scene(b, n, centres, False)
        rep.run(name, thr, lambda: b.get_collisions_many(first)[1])
    b.close()


def bench_bullet_raycast(rep, backend, seed):
    centres = random_circles(100_000, -50.0, 50.0, seed)
    deltas = random_translations(100_000, -5.0, 5.0, seed + 1)
    ray = np.float32([[0.0, -60.0, 0.0, 60.0]])
    b = harness.Backend(backend, NONE)
    for n, label, thr in ((10_000, "10k", 1.0), (100_000, "100k", 10.0)):
        ids = np.arange(n, dtype=np.uint32)

        def reset():
            b.clear()
            b.create_entities(ids, np.full(n, BULLET, np.uint8), xf3(centres[:n]))
            b.add_shapes(ids, np.zeros(n, np.uint8), circle_geom(n, 1.0))
            b.move_transform(ids, xf3(deltas[:n]))
        reset()
        # only the 10k case re-creates the scene before every sample (test_registry_raycast_performance.cpp:34-43)
        rep.run(f"Registry | Bullet raycast among {label} entities", thr, lambda: b.raycast(ray, with_time=True)[2],
                reset if n == 10_000 else None)
    b.close()


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--backend", default="b200", choices=["b200", "reference", "oracle"])
    ap.add_argument("--out", default="bench_out", help="folder of the benchmarks--*.md file (tests/main.cpp:42)")
    ap.add_argument("--samples", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=1, help="untimed calls before the samples (Catch2's warm-up phase)")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--filter", default="", help="substring of the benchmark name (the reference's --filter-benchmark)")
    args = ap.parse_args()
    if not harness.available(args.backend):
        sys.exit(f"backend {args.backend!r} is not built (python colli2de_b200/build.py)")
    rep = Report(args.samples, args.filter, args.warmup)
    for bench in (bench_insert, bench_move, bench_all_pairs, bench_single_queries, bench_bullet_raycast):
        bench(rep, args.backend, args.seed)
    path = rep.export(args.out)
    print(path or "no benchmark matched the filter")


if __name__ == "__main__":
    main()
