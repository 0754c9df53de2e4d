#!/usr/bin/env python
"""bench.py — Colli2De per-frame collision query path on B200: getCollidingPairs ms/frame and pairs/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload cfg3-clustered|cfg3-uniform|cfg2|cfg5-rays] [--scaling strong|weak]

Own arm (default): the BASELINE.json workload quoted by the metric — cfg 3, 1 M dynamic mixed shapes
(circle / capsule / segment / convex polygon) + 100 k static polygons, clustered density — driven through the
drop-in c2d::Registry<uint32_t> (tests/cpp/runner.cpp compiled against include/colli2de/Registry.hpp)
-> C ABI (include/c2d_gpu.h) -> hand-written sm_100a kernels.  One "step" = one frame: every dynamic entity
is moved by a jitter in U(-0.5, 0.5)^2, then getCollidingPairs runs.
  * `value` / `ms_per_step`: device-resident path — the K frames' move batches are staged in HBM before the
    timed region (Registry::stageMoves), each step = applyStagedMoves + the whole pipeline (refit, Morton,
    radix sort, LBVH build + fit, traversal, binning, narrowphase); results stay on the device.
  * `e2e`: the reference-facing call with HOST buffers — Registry::getCollidingPairs() after per-entity
    moveEntity calls: the pinned move log is uploaded (H2D) and the 84-byte EntityCollision records are
    downloaded (D2H) and returned as std::vector inside the timed call.  Like the reference's metric
     the per-entity moveEntity host calls are outside the timed call; their cost is
    reported as e2e.moves_host_ms.
  * `parity`: before the timed regions two frames are compared with the UNMODIFIED reference (oracle/_ref) on rank 0 —
    sorted tuple set, manifolds, orientation-ambiguous pairs classified (tests/parity.py).  At N > 1 the records of all
    ranks are gathered to rank 0 by the library (C2D_RESULTS_ROOT) for it.
  * `roofline`: the dominant kernel (chosen from warm-up frames profiled with CUDA events around every kernel), its
    average launch duration measured with CUDA events around each of its launches INSIDE the timed region
    (c2d_gpu_set_profiling + c2d_gpu_set_profile_filter: events around every kernel would cost 0.12 ms per frame),
    algorithmic bytes per launch as defined in DESIGN.md section 3, against MEASURED_PEAKS.json hbm_gbs; `traffic` from the
    newest profiles/*_ncu_traffic.json.
  * `cpu_baseline`: the UNMODIFIED reference (oracle/_ref, compiled from /root/reference by oracle/Makefile)
    timed on this host on a bounded sample (same scene, a few frames).
Reference arm (--impl reference): the reference's own CPU getCollidingPairs on the same scene, K steps.

Multi-GPU (torchrun, one rank per GPU): the decomposition lives in the library (include/c2d_gpu.h "multi-GPU").  Every
rank holds a Registry of the whole scene joined to one NCCL group (Registry::joinGroup); each frame a rank refits and
traverses only its slab of the Morton order of the Dynamic tree, lets the halo shapes of higher ranks query it, and
returns its share of the pairs.  Default `scaling` is "strong": the metric's fixed 1 M-shape scene on N GPUs.
The same invocation then measures a second line, `weak`: an N-times larger world (N x the scene, mirrored on every rank)
through the same library path, device-resident loop only (`--no-weak-line` skips it; `--scaling weak` runs the whole
bench on the larger world instead).  In the e2e loop every rank calls moveEntity for 1/N of the entities and
the library exchanges the 12-byte records over NVLink (Registry::setGroupMoveExchange) inside getCollidingPairs.

--workload cfg5-rays: 1 M batched rays (half finite, half infinite) against a 4 M-shape scene, rays/s; device-resident =
c2d_gpu_raycast_begin (hits stay in HBM), e2e = Registry::rayCastBatch with host buffers; N > 1 = ray sharding.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

from colli2de_b200 import scenes  # noqa: E402

WORKLOADS = {
    # name: (n_dynamic, n_static, box, clustered)
    "cfg3-clustered": (1_000_000, 100_000, 5000.0, True),
    "cfg3-uniform": (1_000_000, 100_000, 5000.0, False),
    "cfg2": (100_000, 10_000, 1600.0, False),
}
METRIC = "getCollidingPairs pairs/sec at 1M shapes (ms/frame = ms_per_step)"
UNIT = "pairs/s"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, flag in zip(names, r[3:7]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_scene(name: str, scale: int = 1):
    """The workload; scale > 1 builds the scale-times larger world of the weak-scaling runs (same density)."""
    n_dyn, n_sta, box, clustered = WORKLOADS[name]
    return (scenes.mixed_scene(n_dyn * scale, n_sta * scale, box * scale ** 0.5, clustered=clustered),
            n_dyn * scale, n_sta * scale)


def measured_peak_gbs():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def geometry_bytes(scene, mask) -> float:
    """mean geometry bytes per shape as in SURVEY.md section 8(d): circle 12, capsule 20, segment 16, polygon 4 + 8*count"""
    t, c = scene["shp_type"][mask], scene["shp_count"][mask].astype(np.float64)
    return float(np.where(t == 0, 12, np.where(t == 1, 20, np.where(t == 2, 16, 4 + 8 * c))).mean())


# ---------------------------------------------------------------------------------------------------------
def run_reference(args, scene, n_dyn, as_baseline: bool):
    """The reference's own CPU getCollidingPairs (oracle/_ref).  as_baseline: bounded sample for cpu_baseline."""
    from tests import harness as H

    which = "reference" if H.available("reference") else "oracle"
    kind = "reference" if which == "reference" else "port"
    b = H.Backend(which)
    scenes.populate(b, scene)
    rng = scenes.LCG(2024)
    dyn = np.arange(n_dyn, dtype=np.uint32)
    steps, warm = (3, 1) if as_baseline else (args.steps, args.warmup)
    # bound the arm: probe one frame, shrink the step count if K frames would take more than ~3 minutes
    times, pairs = [], 0
    budget_s = 25.0 if as_baseline else 170.0
    t_start = time.perf_counter()
    for f in range(warm + steps):
        b.move_translate(dyn, scenes.jitter(rng, n_dyn, 0.5))
        sec = C.c_double(0.0)
        n = int(b.lib.rn_get_colliding_pairs(b.ctx, C.byref(sec)))
        b._check()
        if f >= warm:
            times.append(sec.value)
            pairs += n
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    done = len(times)
    total = float(np.sum(times))
    threads = 5 if len(scene["shp_entity"]) > 9000 else 1  # Registry.hpp:469,536-539: 5 std::threads, narrowphase serial
    return dict(value=pairs / total, unit=UNIT, cores=threads, kind=kind, steps=done, ms_per_step=1e3 * total / done,
                pairs_per_frame=pairs / done, host_cores=os.cpu_count(),
                sample=(f"full workload scene, {done} timed frames after {warm} warm-up; reference Release flags "
                        f"(-O3 -DNDEBUG -ffp-contract=off), broadphase on <= {threads} std::threads, narrowphase serial; "
                        f"steady_clock around getCollidingPairs() only"))


# ---------------------------------------------------------------------------------------------------------
def workload_label(name: str, n_dyn: int, n_sta: int) -> str:
    """config.workload, the same string on both arms"""
    return (f"{name}: {n_dyn} dynamic mixed circle/capsule/segment/polygon + {n_sta} static polygons, per-frame jitter "
            f"U(-0.5,0.5)^2, LCG seed 4242 (SURVEY.md 8(d) cfg 3)")


def parity_frames(b, scene, dyn, chunk, rank, world, frames: int):
    """`frames` frames of the e2e configuration (move exchange on, results gathered to rank 0 when N > 1) compared with
    the UNMODIFIED reference on rank 0: exact tuple set, manifolds within 1e-5 (bit-identical counted), orientation-
    ambiguous segment pairs classified.  Collective over the ranks."""
    from tests import harness as H, parity as P

    which = "reference" if H.available("reference") else "oracle"
    ref, idx = None, None
    if rank == 0:
        ref = H.Backend(which)
        scenes.populate(ref, scene)
        idx = P.SceneIndex(scene)
    rng = scenes.LCG(777)
    out = {"checker": "oracle/_ref (the unmodified reference)" if which == "reference" else "oracle/c2d_oracle.c",
           "frames": frames, "expected": 0, "got": 0, "ambiguous": 0, "bit_identical": 0, "not_bit_identical": 0, "ok": True}
    for _ in range(frames):
        d = scenes.jitter(rng, len(dyn), 0.5)
        b.move_translate(dyn[chunk], d[chunk])
        got = b.get_colliding_pairs()
        if rank == 0:
            ref.move_translate(dyn, d)
            try:
                res = P.compare_pairs(which, ref, idx, ref.get_colliding_pairs(), got)
                for k in ("expected", "got", "ambiguous", "bit_identical", "not_bit_identical"):
                    out[k] += res[k]
            except AssertionError as e:  # reported, not hidden: the line says parity failed
                out["ok"] = False
                out["error"] = str(e)[:300]
    if ref is not None:
        ref.close()
    return out


def run_b200(args, scene, n_dyn, n_sta, rank, world, local, weak):
    import torch
    import torch.distributed as dist

    from colli2de_b200 import capi
    from tests import harness as H

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (own arm) needs a CUDA device: colli2de_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    capi.lib()  # fail loudly if libc2d_gpu.so is missing
    b = H.Backend("b200", 1 if args.grid_cell > 0 else 0, max(args.grid_cell, 1))
    scenes.populate(b, scene)  # every rank mirrors the scene; the per-frame work is what the library shards
    dyn = np.arange(n_dyn, dtype=np.uint32)
    ctx = b.native_context()
    if world > 1:
        b.group_join_torch()  # Registry::joinGroup: one NCCL communicator over the ranks' contexts
    capi.set_rebuild_period(ctx, args.rebuild_period)
    K, W = args.steps, args.warmup
    # the entities this rank drives in the end-to-end loop (move exchange): one contiguous block of the dynamic ids
    chunk = slice((n_dyn * rank) // world, (n_dyn * (rank + 1)) // world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- parity against the reference, in the end-to-end configuration ---------------------------------------------
    if world > 1:
        b.group_options(results=capi.RESULTS_ROOT, exchange=True)
    parity = None
    if args.parity_frames > 0:
        parity = parity_frames(b, scene, dyn, chunk if world > 1 else slice(None), rank, world, args.parity_frames)
    if world > 1:
        b.group_options(results=capi.RESULTS_LOCAL, exchange=False)

    # ---- device-resident arm: move batches staged in HBM before the timed region -------------------------------
    rng = scenes.LCG(2024)

    def frame_moves():
        return scenes.jitter(rng, n_dyn, 0.5)

    # Warm-up with per-kernel CUDA events on ALL kernels: the full kernel table, and which kernel dominates.  The events
    # cost ~0.12 ms per frame (one pair per launch, measured: --probe-profiling-cost), so in the TIMED region only the
    # dominant kernel keeps its event pair (free) - its average launch duration there is what the roofline divides by.
    handles = [b.stage_translations(dyn, frame_moves()) for _ in range(2 * W + K)]
    all_handles = handles
    for h in handles[:W]:
        b.apply_staged(h)
        b.count_pairs_device()
    capi.set_profiling(ctx, True)
    capi.kernel_profile(ctx, reset=True)
    for h in handles[W:2 * W]:
        b.apply_staged(h)
        b.count_pairs_device()
    warm_profile = capi.kernel_profile(ctx, reset=True)
    warm_frames = W
    dom_name = max(warm_profile.items(), key=lambda kv: kv[1][1])[0]
    plain_ms = None
    if args.probe_profiling_cost:
        extra = [b.stage_translations(dyn, frame_moves()) for _ in range(K)]
        barrier()
        tp = time.perf_counter()
        b.run_staged_frames(extra)  # all kernels timed
        barrier()
        plain_ms = 1e3 * allmax(time.perf_counter() - tp) / K
        capi.kernel_profile(ctx, reset=True)
        for h in extra:
            b.release_staged(h)
    capi.set_profile_filter(ctx, dom_name)
    handles = handles[W:]  # the timed loop below skips its first W again
    sampler = ClockSampler(local) if rank == 0 else None
    stage_ms = dict(ms_refit=0.0, ms_build=0.0, ms_broad=0.0, ms_narrow=0.0, ms_device=0.0)
    slab = dict(slab_owned=0.0, slab_leaves=0.0, ms_slab=0.0)
    barrier()
    t0 = time.perf_counter()
    # the K timed frames run in one native loop (tests/cpp/runner.cpp rn_run_staged_frames): Registry::applyStagedMoves +
    # Registry::countCollidingPairsOnDevice per frame, no Python in between
    pairs_total, sums, _ = b.run_staged_frames(handles[W:])
    cand_total = sums["candidates"]
    launches = int(sums["kernel_launches"])
    for k in stage_ms:
        stage_ms[k] = sums[k]
    if world > 1:
        cs = capi.comm_stats(ctx)  # of the last frame
        for k in slab:
            slab[k] = cs[k] * K
    barrier()
    elapsed = allmax(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    profile = capi.kernel_profile(ctx, reset=True)  # the dominant kernel only, over the timed region
    capi.set_profile_filter(ctx, None)
    capi.set_profiling(ctx, False)
    for h in all_handles:
        b.release_staged(h)
    pairs_all = allsum(float(pairs_total))
    cand_all = allsum(float(cand_total))
    per_rank = None
    if world > 1:
        rows = [None] * world
        dist.all_gather_object(rows, {"pairs_per_frame": pairs_total / K, "candidates_per_frame": cand_total / K,
                                      "ms_device": stage_ms["ms_device"] / K, "ms_broad": stage_ms["ms_broad"] / K,
                                      "ms_narrow": stage_ms["ms_narrow"] / K, "ms_build": stage_ms["ms_build"] / K,
                                      "slab_owned": slab["slab_owned"] / K, "slab_halo": (slab["slab_leaves"] - slab["slab_owned"]) / K})
        per_rank = rows
    ms_per_step = 1e3 * elapsed / K
    value = pairs_all / elapsed

    # ---- end-to-end arm: host buffers through the drop-in API ----------------------------------------------------
    # N > 1: every rank calls moveEntity for the block of entities it drives (1/N of the host work and of the upload);
    # the library all-gathers the 12-byte records over NVLink inside getCollidingPairs (move exchange), each rank
    # computes its Morton slab and returns ITS share of the pairs into its own host memory.
    def e2e_loop(frames: int, skip: int):
        times, moves, pairs, up, down, xms, gms, xbytes, gbytes = [], [], 0, 0.0, 0.0, [], [], 0.0, 0.0
        stages = dict(ms_refit=0.0, ms_build=0.0, ms_broad=0.0, ms_narrow=0.0, ms_device=0.0, ms_total=0.0, ms_host_enqueue=0.0,
                      ms_host_wait=0.0)
        for f in range(skip + frames):
            d = frame_moves()
            if f == skip:
                barrier()
            tm = time.perf_counter()
            b.move_translate(dyn[chunk], d[chunk])
            tm = time.perf_counter() - tm
            if world > 1:
                barrier()  # the collective call starts together on every rank: the ranks' host move loops differ by milliseconds
            sec = C.c_double(0.0)
            n = int(b.lib.rn_get_colliding_pairs(b.ctx, C.byref(sec)))
            b._check()
            if f >= skip:
                st = b.stats()
                times.append(sec.value)
                moves.append(tm)
                for k in stages:
                    stages[k] += st[k] / frames
                pairs += n
                up += st["h2d_bytes"]
                down += st["d2h_bytes"]
                if world > 1:
                    cs = capi.comm_stats(ctx)
                    xms.append(cs["ms_exchange"])
                    gms.append(cs["ms_gather"])
                    xbytes += cs["exchange_bytes"]
                    gbytes += cs["gather_bytes"]
        barrier()
        ranks = None
        if world > 1:
            ranks = [None] * world
            dist.all_gather_object(ranks, {"mean_ms": 1e3 * float(np.mean(times)), "max_ms": 1e3 * float(np.max(times)),
                                           "median_ms": 1e3 * float(np.median(times)), "lib_ms_total": stages["ms_total"],
                                           "lib_ms_host_wait": stages["ms_host_wait"], "pairs": pairs / frames})
        return dict(ranks=ranks, median_ms=1e3 * float(np.median(times)), elapsed=allmax(float(np.sum(times))), pairs=pairs, h2d=allsum(up) / frames, d2h=allsum(down) / frames,
                    moves_ms=allmax(1e3 * float(np.mean(moves))), exchange_ms=float(np.mean(xms)) if xms else 0.0,
                    gather_ms=float(np.mean(gms)) if gms else 0.0, exchange_bytes=xbytes / frames, gather_bytes=gbytes / frames,
                    stages=stages)

    if world > 1:
        b.group_options(results=capi.RESULTS_LOCAL, exchange=True)
    e2e = e2e_loop(K, 2)
    e2e_pairs_all = allsum(float(e2e["pairs"]))
    root = None
    if world > 1:
        # the single-caller view: all records gathered to rank 0 over NVLink and downloaded there
        b.group_options(results=capi.RESULTS_ROOT, exchange=True)
        root = e2e_loop(min(K, 8), 1)
        b.group_options(results=capi.RESULTS_LOCAL, exchange=False)
    # N2: the same frame driven through Registry::moveEntities (one call for the whole batch) instead of one moveEntity
    # call per entity
    bulk_move_ms, bulk_call_ms = [], []
    if world > 1:
        b.group_options(results=capi.RESULTS_LOCAL, exchange=True)
    for f in range(1 + min(K, 8)):
        d = frame_moves()
        tb = b.move_entities_bulk(dyn[chunk], d[chunk])
        if world > 1:
            barrier()
        sec = C.c_double(0.0)
        b.lib.rn_get_colliding_pairs(b.ctx, C.byref(sec))
        b._check()
        if f >= 1:
            bulk_move_ms.append(1e3 * tb)
            bulk_call_ms.append(1e3 * sec.value)
    if world > 1:
        b.group_options(results=capi.RESULTS_LOCAL, exchange=False)
    # same loop through the zero-copy extension (a view of the pinned result buffer instead of a fresh std::vector)
    view_times = []
    if world == 1:
        for f in range(1 + min(K, 8)):
            b.move_translate(dyn, frame_moves())
            sec, ptr = C.c_double(0.0), C.c_void_p()
            b.lib.rn_get_colliding_pairs_view(b.ctx, C.byref(sec), C.byref(ptr))
            b._check()
            if f >= 1:
                view_times.append(sec.value)
    barrier()
    e2e_elapsed = e2e["elapsed"]
    if world > 1:
        b.group_leave()

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (live CUDA-event times) ----------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    dom_count, dom_ms = profile.get(dom_name, (0, 0.0))
    if dom_count == 0:  # the kernel did not run in the timed frames (e.g. a rebuild-only kernel in a very short run)
        dom_count, dom_ms = warm_profile[dom_name]
    n_shapes = n_dyn + n_sta
    cand_per_frame = cand_all / K / world        # per rank
    pairs_per_frame = pairs_all / K              # whole job
    n_query = slab["slab_owned"] / K if world > 1 and slab["slab_owned"] else n_dyn   # Dynamic query leaves of this rank
    n_tree = slab["slab_leaves"] / K if world > 1 and slab["slab_leaves"] else n_dyn  # leaves of its Dynamic tree
    g_all = geometry_bytes(scene, np.ones(n_shapes, bool))
    g_np = float(np.where(scene["shp_type"] == 3, 4 + 16 * scene["shp_count"].astype(np.float64),
                          np.where(scene["shp_type"] == 0, 12, np.where(scene["shp_type"] == 1, 20, 16))).mean())
    alg = {
        # per-launch algorithmic bytes of each kernel (DESIGN.md "Roofline")
        "k_apply_translates": n_dyn * (12 + 4 + 16 + 16 + 16 + 16 + 16 + 16),
        "k_traverse<self>": n_query * (4 + 16 + 8) + (n_tree - 1) * 64 + 0.8 * cand_per_frame * 8,
        "k_traverse<cross>": n_query * (4 + 16 + 8) + (n_sta - 1) * 64 + 0.2 * cand_per_frame * 8,
        # narrowphase: the polygon bin (polygon vs polygon|capsule|segment) holds ~45 % of the candidates of this mix
        "k_narrow_filter": 0.55 * cand_per_frame * (8 + 2 * (4 + 4 + 16 + 20)) + 0.2 * cand_per_frame * 16,
        "k_narrow_filter_pp": 0.45 * cand_per_frame * (8 + 2 * (4 + 4 + 16 + 4 + 16 * 5.5)) + 0.15 * cand_per_frame * 16,
        "k_narrow_manifold": 0.55 * pairs_per_frame / world * (16 + 2 * (4 + 4 + 16 + 20) + 84),
        "k_narrow_manifold_pp": 0.45 * pairs_per_frame / world * (16 + 2 * (4 + 4 + 16 + 4 + 16 * 5.5) + 84),
        "radix_sort(1 memset + 5 kernels)": n_tree * (4 + 4 * 16),
        "k_fit": n_tree * (4 + 16 + 8 + 4) + (n_tree - 1) * (64 + 4 + 4),
        "k_karras_build": n_tree * 4 + (n_tree - 1) * (16 + 8),
        "k_morton": n_dyn * (4 + 16 + 8),
        "k_bounds": n_dyn * (4 + 16),
        "k_slab_keys": n_dyn * (4 + 16 + 8),
        "k_slab_hist2": n_dyn * 4,
        "k_slab_mark": n_dyn * (8 + 16 / world),
        "k_slab_select": n_dyn * (8 + 16) + n_tree * 8,
        "k_bin_count": cand_per_frame * (8 + 8),
        "k_bin_scatter": cand_per_frame * (8 + 8 + 8),
    }
    frame_alg = n_shapes * (20 + g_all + 16 + 16 + 8) + cand_per_frame * world * (16 + 2 * (20 + g_np)) + pairs_per_frame * 84
    dom_avg_ms = dom_ms / max(dom_count, 1)
    dom_bytes = float(alg.get(dom_name, 0.0))
    achieved = dom_bytes / (dom_avg_ms * 1e-3) / 1e9 if dom_avg_ms > 0 else 0.0
    kernels = {k: {"launches_per_frame": c / warm_frames, "avg_ms": ms / max(c, 1),
                   "share": ms / max(sum(v[1] for v in warm_profile.values()), 1e-9),
                   "alg_gbs": (alg.get(k, 0.0) / (ms / max(c, 1) * 1e-3) / 1e9) if ms > 0 else 0.0}
               for k, (c, ms) in sorted(warm_profile.items(), key=lambda kv: -kv[1][1])}
    # DRAM traffic per launch of the dominant kernel from the newest committed `ncu --set full` capture (profiles/)
    traffic, traffic_src = None, None
    tpaths = sorted(p for p in os.listdir(os.path.join(REPO, "profiles")) if p.endswith("_ncu_traffic.json"))
    if tpaths and world == 1 and args.workload == "cfg3-clustered":
        tdoc = json.load(open(os.path.join(REPO, "profiles", tpaths[-1])))
        ncu_name = {"k_traverse<self>": "k_traverse<1>", "k_traverse<cross>": "k_traverse<0>"}.get(dom_name, dom_name)
        rec = tdoc.get(ncu_name)
        if rec:
            traffic = rec["dram"]
            traffic_src = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/{tpaths[-1]} "
                           f"({tdoc.get('_source', 'ncu --set full')}; cold-cache, serialised replay; "
                           f"{rec['launches_profiled']} launches)")
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "alg_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_avg_ms,
                "frame": {"alg_bytes": frame_alg, "achieved": frame_alg / (ms_per_step * 1e-3) / 1e9,
                          "frac": frame_alg / (ms_per_step * 1e-3) / 1e9 / peak},
                "note": "latency/divergence-bound path (SURVEY.md 8(d)): whole-frame algorithmic traffic is ~0.2-0.3 GB",
                "timing": f"CUDA events around every launch of {dom_name} inside the timed region ({dom_count} launches); the table "
                          f"`kernels` comes from {warm_frames} warm-up frames with events around every kernel (kept out of the timed "
                          f"region: one event pair per launch costs ~0.12 ms per frame)",
                "kernels": kernels}

    if world == 1:
        parallelism = "single GPU"
    else:
        parallelism = (f"{'weak: ' + str(world) + ' x the workload in one world' if weak else 'strong: the fixed workload'} on {world} "
                       f"GPUs, one process each; in-library decomposition (include/c2d_gpu.h multi-GPU): scene mirrored, the "
                       f"Dynamic tree cut into Morton-key slabs balanced from the key histogram every frame (rank 0: "
                       f"{n_query:.0f} owned + {n_tree - n_query:.0f} halo leaves of {n_dyn}), owner rule, no data-path "
                       f"collective in the device-resident loop; e2e loop: the 12-byte move records of every rank broadcast over NVLink")
    e2e_out = {"value": e2e_pairs_all / e2e_elapsed, "unit": UNIT, "ms_per_step": 1e3 * e2e_elapsed / K,
               "median_ms_per_step": (max(r["median_ms"] for r in e2e["ranks"]) if e2e["ranks"] else e2e["median_ms"]),
               "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
               "moves_host_ms": e2e["moves_ms"],
               "whole_frame_ms": e2e["moves_ms"] + 1e3 * e2e_elapsed / K,
               "rank0_stage_ms": e2e["stages"],
               "bulk_moves": {"call": "Registry::moveEntities(ids, deltas): one call per frame instead of one moveEntity per entity",
                              "moves_host_ms": float(np.mean(bulk_move_ms)),
                              "query_ms": float(np.mean(bulk_call_ms)),
                              "whole_frame_ms": float(np.mean(bulk_move_ms)) + float(np.mean(bulk_call_ms))},
               "call": ("c2d::Registry<uint32_t>::getCollidingPairs() after moveEntity(id, Translation) per entity" if world == 1
                        else "per rank: moveEntity(id, Translation) for the 1/N of the entities it drives, then the collective "
                             "c2d::Registry<uint32_t>::getCollidingPairs() returning the rank's share into host memory; time = "
                             "max over ranks, pairs = sum over ranks")}
    if world == 1:
        e2e_out["view_ms_per_step"] = 1e3 * float(np.mean(view_times))
        e2e_out["view_call"] = "Registry::getCollidingPairsView() (extension: span over the pinned result buffer)"
    else:
        e2e_out["per_rank_call_ms"] = e2e["ranks"]
        e2e_out["collective"] = {"name": "ncclAllGather of a 16-byte header + one grouped ncclBroadcast per rank of its 12-byte move records (device to device)",
                                 "bytes_received_per_rank_per_step": e2e["exchange_bytes"],
                                 "device_ms_per_step": e2e["exchange_ms"]}
        e2e_out["root_gather"] = {"what": "GroupResults::Root - all records gathered to rank 0 over NVLink (ncclAllGather of "
                                          "the counts + ncclSend/ncclRecv) and returned by rank 0's getCollidingPairs()",
                                  "ms_per_step": 1e3 * root["elapsed"] / min(K, 8),
                                  "gather_bytes_per_step": root["gather_bytes"], "gather_device_ms": root["gather_ms"]}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args.workload, n_dyn, n_sta),
                   "parallelism": parallelism,
                   "l2": "inputs larger than L2 (fresh 12 MB move batch per frame; ~255 MB shape state + ~100 MB "
                         "sort/tree buffers rewritten every frame vs 126 MB L2), no explicit flush",
                   "lbvh_rebuild_period": args.rebuild_period,
                   "broadphase": f"uniform grid, cell {args.grid_cell}" if args.grid_cell > 0 else "LBVH",
                   "pairs_per_frame": pairs_per_frame, "candidates_per_frame": cand_per_frame * world,
                   "stage_ms_per_step": {k: v / K for k, v in stage_ms.items()},
                   "ms_per_step_with_events_on_every_kernel": plain_ms},
        "e2e": e2e_out,
        "parity": parity,
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
    }
    if world > 1:
        out["config"]["slab_ms_per_step"] = slab["ms_slab"] / K
        out["config"]["per_rank"] = per_rank
    return out


def run_weak_line(args, rank, world, local):
    """Second line of a multi-GPU run: the same in-library decomposition on an N-times larger world
    (weak scaling), device-resident loop only.  Returns the sub-object `weak` of the JSON line (rank 0) or None."""
    import torch
    import torch.distributed as dist

    from colli2de_b200 import capi
    from tests import harness as H

    scene, n_dyn, n_sta = build_scene(args.workload, world)
    b = H.Backend("b200")
    scenes.populate(b, scene)
    dyn = np.arange(n_dyn, dtype=np.uint32)
    ctx = b.native_context()
    b.group_join_torch()
    capi.set_rebuild_period(ctx, args.rebuild_period)
    K, W = min(args.steps, 10), max(args.warmup, 3)
    rng = scenes.LCG(2024)
    handles = [b.stage_translations(dyn, scenes.jitter(rng, n_dyn, 0.5)) for _ in range(W + K)]
    for h in handles[:W]:
        b.apply_staged(h)
        b.count_pairs_device()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pairs, sums, _ = b.run_staged_frames(handles[W:])
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0, float(pairs), -float(pairs)], dtype=torch.float64, device="cuda")
    mx = t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    elapsed, pairs_all, pairs_max, pairs_min = float(mx[0]), float(t[1]), float(mx[1]), -float(mx[2])
    for h in handles:
        b.release_staged(h)
    b.group_leave()
    b.close()
    if rank != 0:
        return None
    return {"scaling": "weak", "value": pairs_all / elapsed, "unit": UNIT, "ms_per_step": 1e3 * elapsed / K, "steps": K, "warmup": W,
            "workload": workload_label(args.workload, n_dyn, n_sta) + f" = {world} x the metric's scene in one world, mirrored on every rank",
            "pairs_per_frame": pairs_all / K, "pairs_per_frame_per_rank_min_max": [pairs_min / K, pairs_max / K],
            "rank0_ms_device": sums["ms_device"] / K,
            "note": "device-resident loop only (moves staged in HBM); same library path as the strong line above"}


# ---------------------------------------------------------------------------------------------------------
# cfg 5: batched rayCast, 1 M rays (odd: finite, length 50; even: infinite) against a 4 M-shape scene
RAY_METRIC = "batched rayCast rays/sec, 1M rays vs 4M shapes (ms/batch = ms_per_step)"
RAY_UNIT = "rays/s"
RAY_SCENE = (3_640_000, 360_000, 9535.0, 77)  # dynamic, static, box, seed (SURVEY.md 8(d) cfg 5: cfg-3 mix, same density)


def ray_workload_label(n_rays: int) -> str:
    return (f"cfg5-rays: {n_rays} rays (even index infinite unit direction, odd index finite length 50, mask ~0) vs "
            f"{RAY_SCENE[0]} dynamic mixed + {RAY_SCENE[1]} static polygons in [0,{RAY_SCENE[2]:.0f}]^2 (SURVEY.md 8(d) cfg 5)")


def ray_reference(n_sample: int, ray4, inf, scene):
    from tests import harness as H

    which = "reference" if H.available("reference") else "oracle"
    r = H.Backend(which)
    scenes.populate(r, scene)
    off, hits, sec = r.raycast(ray4[:n_sample], inf[:n_sample], with_time=True)
    return r, which, off, hits, sec


def run_rays_reference(args):
    scene = scenes.mixed_scene(RAY_SCENE[0], RAY_SCENE[1], RAY_SCENE[2], seed=RAY_SCENE[3])
    ray4, inf = scenes.rays(args.rays, RAY_SCENE[2], seed=13)
    m = min(args.rays, 4000)
    _, which, off, hits, sec = ray_reference(m, ray4, inf, scene)
    return {"impl": "reference", "metric": RAY_METRIC, "value": m / sec, "unit": RAY_UNIT, "n_gpus": args.gpus, "steps": 1,
            "warmup": 0, "ms_per_step": 1e3 * sec * args.rays / m, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": ray_workload_label(args.rays)},
            "cpu_baseline": {"value": m / sec, "unit": RAY_UNIT, "cores": 1, "kind": "reference" if which == "reference" else "port",
                             "sample": f"the first {m} rays of the batch looped through Registry::rayCast(ray) (single thread); "
                                       f"ms_per_step is the extrapolation to the whole batch", "host_cores": os.cpu_count()},
            "e2e": {"value": m / sec, "unit": RAY_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def run_rays(args, rank, world, local):
    import torch
    import torch.distributed as dist

    from colli2de_b200 import capi
    from tests import harness as H

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (own arm) needs a CUDA device: colli2de_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.lib()
    scene = scenes.mixed_scene(RAY_SCENE[0], RAY_SCENE[1], RAY_SCENE[2], seed=RAY_SCENE[3])
    n = args.rays
    ray4, inf = scenes.rays(n, RAY_SCENE[2], seed=13)
    b = H.Backend("b200")
    scenes.populate(b, scene)  # every rank mirrors the scene; the ray batch is what is sharded
    ctx = b.native_context()
    if world > 1:
        capi.set_shard(ctx, rank, world)
        b.group_ray_sharding(True)
    K, W = args.steps, max(args.warmup, 3)
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allred(x: float, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    # ---- parity: the first rays of the batch against the reference's rayCast on rank 0 ------------------------------
    parity = None
    if args.parity_frames > 0:
        m = min(n, 1500)
        off, hits = b.raycast(ray4[:m], inf[:m])
        counts = (off[1:] - off[:-1]).astype(np.int64)
        parts = [(counts.tobytes(), hits.tobytes())]
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, (counts.tobytes(), hits.tobytes()))
        if rank == 0:
            got_counts = sum(np.frombuffer(c, np.int64) for c, _ in parts)  # every ray is non-empty on one rank only
            got_hits = np.concatenate([np.frombuffer(h, hits.dtype) for _, h in parts])  # rank order = ray order
            got_off = np.concatenate([[0], np.cumsum(got_counts)])
            ref, which, roff, rhits, _ = ray_reference(m, ray4, inf, scene)
            same = sum(1 for i in range(m) if rhits[int(roff[i]):int(roff[i + 1])].tobytes() ==
                       got_hits[int(got_off[i]):int(got_off[i + 1])].tobytes())
            parity = {"checker": "oracle/_ref (the unmodified reference)" if which == "reference" else "oracle/c2d_oracle.c",
                      "rays": m, "rays_identical": same, "hits_expected": int(len(rhits)), "hits_got": int(len(got_hits)),
                      "note": "rays whose hit list differs hold hits with exactly equal (entry, exit) keys, which std::set keeps "
                              "one of by traversal order (SURVEY.md D10)", "ok": same >= m - 8 and len(rhits) == len(got_hits)}
            ref.close()

    # ---- device-resident arm: c2d_gpu_raycast_begin (the hits stay in HBM; offsets and the total come back) -----------
    RAY = np.dtype([("sx", "<f4"), ("sy", "<f4"), ("ex", "<f4"), ("ey", "<f4"), ("mask", "<u8"), ("infinite", "<u4"), ("pad", "<u4")])
    rays = np.zeros(n, RAY)
    rays["sx"], rays["sy"], rays["ex"], rays["ey"] = ray4[:, 0], ray4[:, 1], ray4[:, 2], ray4[:, 3]
    rays["mask"], rays["infinite"] = np.uint64(0xFFFFFFFFFFFFFFFF), inf
    stats = capi.Stats()
    lib.c2d_gpu_raycast_begin.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p),
                                          C.POINTER(capi.Stats)]
    total, offs = C.c_uint64(0), C.c_void_p()

    def begin():
        rc = lib.c2d_gpu_raycast_begin(ctx, n, rays.ctypes.data, 0, C.byref(total), C.byref(offs), C.byref(stats))
        if rc != 0:
            raise RuntimeError(lib.c2d_gpu_last_error(ctx).decode())
        return int(total.value)

    for _ in range(W):
        begin()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t0 = time.perf_counter()
    hits_total, dev_ms, launches, leaf_hits = 0, 0.0, 0, 0
    for _ in range(K):
        hits_total += begin()
        dev_ms += stats.ms_device
        launches += int(stats.kernel_launches)
        leaf_hits += int(stats.candidates)
    barrier()
    elapsed = allred(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
    clocks = sampler.stop() if sampler else None
    hits_all = allred(float(hits_total), dist.ReduceOp.SUM if world > 1 else None)
    leaf_all = allred(float(leaf_hits), dist.ReduceOp.SUM if world > 1 else None)
    dev_ms_max = allred(dev_ms, dist.ReduceOp.MAX if world > 1 else None)

    # ---- end-to-end arm: Registry::rayCastBatch with host buffers (rays up, hit records down into a std::vector) ----------
    offsets = np.zeros(n + 1, np.uint64)
    e2e_times, up, down = [], 0.0, 0.0
    for f in range(1 + K):
        if f == 1:
            barrier()
        if world > 1:
            barrier()
        sec = C.c_double(0.0)
        b.lib.rn_raycast(b.ctx, n, ray4.ctypes.data_as(C.POINTER(C.c_float)), inf.ctypes.data_as(C.POINTER(C.c_uint8)), None,
                         offsets.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(sec))
        b._check()
        if f >= 1:
            st = b.stats()
            e2e_times.append(sec.value)
            up += st["h2d_bytes"]
            down += st["d2h_bytes"]
    barrier()
    e2e_elapsed = allred(float(np.sum(e2e_times)), dist.ReduceOp.MAX if world > 1 else None)
    up_all = allred(up, dist.ReduceOp.SUM if world > 1 else None)
    down_all = allred(down, dist.ReduceOp.SUM if world > 1 else None)
    if rank != 0:
        return None
    peak, peak_src = measured_peak_gbs()
    g_np = float(np.where(scene["shp_type"] == 3, 4 + 16 * scene["shp_count"].astype(np.float64),
                          np.where(scene["shp_type"] == 0, 12, np.where(scene["shp_type"] == 1, 20, 16))).mean())
    # bytes_alg(rayCast) = R*24 + L*(16 + 8) + H_c*(20 + g_np) + H*32, per batch and per rank
    R, L, Hh = (hi - lo), leaf_all / K / world, hits_all / K / world
    alg = R * 24 + L * 24 + L * (20 + g_np) + Hh * 32
    batch_ms = dev_ms_max / K
    achieved = alg / (batch_ms * 1e-3) / 1e9
    return {
        "metric": RAY_METRIC, "value": n * K / elapsed, "unit": RAY_UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": ray_workload_label(n),
                   "parallelism": "single GPU" if world == 1 else
                   f"ray sharding (SURVEY.md 8(e)): the scene is mirrored on {world} GPUs, every rank is handed the whole batch "
                   f"and casts its contiguous share ({hi - lo} rays); no exchange, every rank downloads its hits over its own PCIe link",
                   "l2": "4 M-shape scene (~1 GB of shape state + trees) and ~8 GB of hit records per batch vs 126 MB L2, no explicit flush",
                   "hits_per_batch": hits_all / K, "leaf_box_hits_per_batch": leaf_all / K},
        "e2e": {"value": n * K / e2e_elapsed, "unit": RAY_UNIT, "ms_per_step": 1e3 * e2e_elapsed / K,
                "h2d_bytes_per_step": up_all / K, "d2h_bytes_per_step": down_all / K,
                "call": "c2d::Registry<uint32_t>::rayCastBatch(span<RayQuery>) -> offsets + std::vector<RayHit> in host memory"
                        + ("" if world == 1 else " (per rank: its share of the batch; time = max over ranks)")},
        "parity": parity, "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "rayCast batch (k_ray_block tiers + k_heavy_* + k_header_gather), whole batch on one rank",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "alg_bytes_per_launch": alg, "avg_launch_ms": batch_ms,
                     "note": "divergent tree walks + per-ray sorting: latency bound, not bandwidth bound"},
    }


JSON_OUT = sys.stdout


def main():
    # stdout carries exactly ONE line, the JSON result: everything else that writes to fd 1 during the run (NCCL prints its
    # version banner there, libraries may print warnings) is sent to stderr
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3-clustered", choices=sorted(WORKLOADS) + ["cfg5-rays"])
    ap.add_argument("--rays", type=int, default=1_000_000, help="cfg5-rays: rays per batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--probe-profiling-cost", action="store_true",
                    help="also time K frames with CUDA events around EVERY kernel (config.ms_per_step_with_events_on_every_kernel)")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="multi-GPU mode: strong (default) = the metric's fixed workload on N GPUs; weak = an N-times larger "
                         "world at the same density.  Either way the library cuts the Dynamic tree into Morton-key slabs")
    ap.add_argument("--no-weak-line", action="store_true",
                    help="N > 1, strong scaling: skip the second measurement (`weak`: an N-times larger world)")
    ap.add_argument("--parity-frames", type=int, default=2,
                    help="frames compared with the unmodified reference on rank 0 before the timed regions (0 = skip)")
    ap.add_argument("--grid-cell", type=int, default=0,
                    help="> 0: use Registry<Grid>(cell) = the uniform-grid sort-and-sweep broadphase instead of the LBVH")
    ap.add_argument("--rebuild-period", type=int, default=8,
                    help="LBVH topology rebuilt every N-th frame, bottom-up refit in between (1 = every frame)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, world, local = dist_env()

    if args.workload == "cfg5-rays":
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps(run_rays_reference(args)), file=JSON_OUT, flush=True)
            return 0
        out = run_rays(args, rank, world, local)
        if rank == 0:
            if world == 1 and not args.no_cpu_baseline:
                out["cpu_baseline"] = run_rays_reference(args)["cpu_baseline"]
            print(json.dumps(out), file=JSON_OUT, flush=True)
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()
        return 0
    if args.impl == "reference":
        if rank != 0:
            return 0
        weak = args.gpus > 1 and args.scaling == "weak"
        scene, n_dyn, n_sta = build_scene(args.workload, args.gpus if weak else 1)
        r = run_reference(args, scene, n_dyn, as_baseline=False)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_label(args.workload, n_dyn, n_sta), "pairs_per_frame": r["pairs_per_frame"]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), file=JSON_OUT, flush=True)
        return 0

    weak = world > 1 and args.scaling == "weak"
    scene, n_dyn, n_sta = build_scene(args.workload, world if weak else 1)
    out = run_b200(args, scene, n_dyn, n_sta, rank, world, local, weak)
    if world > 1 and not weak and not args.no_weak_line:
        del scene
        second = run_weak_line(args, rank, world, local)
        if rank == 0:
            out["weak"] = second
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            r = run_reference(args, scene, n_dyn, as_baseline=True)
            out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores")}
            out["cpu_baseline"]["ms_per_step"] = r["ms_per_step"]
        print(json.dumps(out), file=JSON_OUT, flush=True)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
