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
  * `roofline`: the dominant kernel by live CUDA-event time (c2d_gpu_set_profiling), algorithmic bytes per
    launch as defined in DESIGN.md section "Roofline", against MEASURED_PEAKS.json hbm_gbs.
  * `cpu_baseline`: the UNMODIFIED reference (oracle/_ref, compiled from /root/reference by oracle/Makefile)
    timed on this host on a bounded sample (same scene, a few frames).
Reference arm (--impl reference): the reference's own CPU getCollidingPairs on the same scene, K steps.

Multi-GPU (torchrun, one rank per GPU): the decomposition lives in the library (include/c2d_gpu.h "multi-GPU").  Every
rank holds a Registry of the whole scene joined to one NCCL group (c2d_gpu_comm_init); each frame a rank sorts, builds
and traverses only its Morton-key slab of the Dynamic tree (+ halo) and returns its share of the pairs.  Default
`scaling` is "strong": the metric's fixed 1 M-shape scene on N GPUs.  `--scaling weak` runs an N-times larger world.
In the e2e loop every rank logs the moves of 1/N of the entities and the library all-gathers the 12-byte records over
NVLink (c2d_gpu_comm_set_move_exchange) before it replays them.
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
def run_b200(args, scene, n_dyn, n_sta, rank, world, local):
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
    weak = world > 1 and args.scaling in ("auto", "weak")
    n_total_dyn, n_total_sta = n_dyn, n_sta
    if weak:
        # this rank's spatial slab of the N-times larger world + halo copies of the neighbours' boundary shapes
        world_body = scene["ent_body"]
        world_ids = scene["ent_id"]
        scene, ghost_mask, local_index, owned_mask, export_mask = local_slab(scene, rank, world)
        owned_dyn = world_ids[owned_mask & (world_body != 0)]
        export_dyn = world_ids[export_mask & (world_body != 0)]
        ghost_dyn = np.sort(scene["ent_id"][ghost_mask & (scene["ent_body"] != 0)])
        ids = scenes.populate(b, scene)
        b.set_ghost(ids[ghost_mask], np.ones(int(ghost_mask.sum()), np.uint8))
        dyn = scene["ent_id"][scene["ent_body"] != 0]
        n_dyn, n_sta = len(dyn), len(scene["ent_id"]) - len(dyn)
    else:
        scenes.populate(b, scene)
        dyn = np.arange(n_dyn, dtype=np.uint32)
    ctx = b.native_context()
    if not weak:
        capi.set_shard(ctx, rank, world)
    capi.set_rebuild_period(ctx, args.rebuild_period)
    K, W = args.steps, args.warmup

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

    # ---- device-resident arm: move batches staged in HBM before the timed region -------------------------------
    rng = scenes.LCG(2024)

    def frame_moves():
        """The frame's move of every dynamic entity this rank holds.  Weak mode: the moves are a function of the global
        entity id (every rank draws the world's stream and keeps its slab), so halo copies move exactly like their
        originals without any exchange."""
        if not weak:
            return scenes.jitter(rng, n_dyn, 0.5)
        return scenes.jitter(rng, n_total_dyn, 0.5)[dyn]

    handles = [b.stage_translations(dyn, frame_moves()) for _ in range(W + K)]
    for h in handles[:W]:
        b.apply_staged(h)
        b.count_pairs_device()
    capi.set_profiling(ctx, True)
    capi.kernel_profile(ctx, reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    launches = 0
    pairs_total = 0
    stage_ms = dict(ms_refit=0.0, ms_build=0.0, ms_broad=0.0, ms_narrow=0.0, ms_device=0.0)
    cand_total = 0
    barrier()
    t0 = time.perf_counter()
    for h in handles[W:]:
        b.apply_staged(h)
        n, _ = b.count_pairs_device()
        st = b.stats()
        pairs_total += n
        cand_total += st["candidates"]
        launches += int(st["kernel_launches"])
        for k in stage_ms:
            stage_ms[k] += st[k]
    barrier()
    elapsed = allmax(time.perf_counter() - t0)
    clocks = sampler.stop() if sampler else None
    profile = capi.kernel_profile(ctx, reset=True)
    capi.set_profiling(ctx, False)
    for h in handles:
        b.release_staged(h)
    pairs_all = allsum(float(pairs_total))
    cand_all = allsum(float(cand_total))
    ms_per_step = 1e3 * elapsed / K
    value = pairs_all / elapsed

    # ---- end-to-end arm: host buffers through the drop-in API ----------------------------------------------------
    from colli2de_b200 import sharding
    from tests.harness import PAIR_DTYPE

    e2e_times, move_times, e2e_pairs, h2d, d2h, gather_times, exchange_times = [], [], 0, 0.0, 0.0, [], []
    pinned = {}
    plan, ghost_ids = None, None
    for f in range(2 + K):
        tx = 0.0
        if weak:
            # the exchange step of the decomposition: a rank decides the moves of the entities it OWNS; the moves of its
            # boundary entities reach the neighbours' halo copies through one NCCL all-gather per frame
            d_world = scenes.jitter(rng, n_total_dyn, 0.5)
            if plan is None:
                # built once for the decomposition: where in the all-gathered buffer each halo copy's move arrives
                plan = sharding.ExchangePlan(export_dyn, ghost_dyn)
                ghost_ids = ghost_dyn[plan.found]
            d_export = d_world[export_dyn]
            dist.barrier()
            tx = time.perf_counter()
            d_ghost = plan.exchange(d_export)
            tx = time.perf_counter() - tx
            step_ids = np.concatenate([owned_dyn, ghost_ids])
            d = np.concatenate([d_world[owned_dyn], d_ghost[plan.found]])
        else:
            step_ids, d = dyn, frame_moves()
        if f == 2:
            barrier()
        tm = time.perf_counter()
        b.move_translate(step_ids, d)
        tm = time.perf_counter() - tm
        sec = C.c_double(0.0)
        tg = 0.0
        # every rank: Registry::getCollidingPairs() of ITS registry (its slab), host vector and all; the union of the
        # ranks' vectors is the world's pair set, each pair exactly once (owner rule; tests/test_gpu_sharding.py)
        n = int(b.lib.rn_get_colliding_pairs(b.ctx, C.byref(sec)))
        b._check()
        st_frame = b.stats()
        if world > 1:
            # optional global view, NOT part of e2e: the 84-byte records of all ranks gathered GPU-to-GPU (NCCL) and
            # downloaded once on rank 0
            dist.barrier()
            tg = time.perf_counter()
            ptr, cnt = capi.device_records(ctx)
            gathered, total = sharding.gather_device_records(ptr, cnt, PAIR_DTYPE.itemsize, dst=0, pinned=pinned)
            torch.cuda.synchronize()
            tg = time.perf_counter() - tg
            if rank == 0 and f == 2:
                recs = np.frombuffer(gathered.tobytes(), dtype=PAIR_DTYPE)
                key = ["entityA", "entityB"] if weak else ["shapeA", "shapeB"]  # shape ids are rank-local in weak mode
                assert len(recs) == total and len(np.unique(recs[key])) == total, "duplicate pair across slabs"
        if f >= 2:
            st = st_frame
            e2e_times.append(sec.value + tx)
            gather_times.append(tg)
            exchange_times.append(tx)
            move_times.append(tm)
            e2e_pairs += n
            h2d += st["h2d_bytes"]
            d2h += st["d2h_bytes"]
    # same loop through the zero-copy extension (a view of the pinned result buffer instead of a fresh std::vector)
    view_times = []
    for f in range(1 + min(K, 8)):
        b.move_translate(dyn, frame_moves())
        sec, ptr = C.c_double(0.0), C.c_void_p()
        b.lib.rn_get_colliding_pairs_view(b.ctx, C.byref(sec), C.byref(ptr))
        b._check()
        if f >= 1:
            view_times.append(sec.value)
    barrier()
    e2e_elapsed = allmax(float(np.sum(e2e_times)))
    e2e_pairs_all = allsum(float(e2e_pairs))

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel (live CUDA-event times) ----------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    dom_name, (dom_count, dom_ms) = max(profile.items(), key=lambda kv: kv[1][1])
    n_shapes = n_dyn + n_sta
    cand_per_frame = cand_all / K / world
    pairs_per_frame = pairs_all / K
    g_all = geometry_bytes(scene, np.ones(n_shapes, bool))
    g_np = float(np.where(scene["shp_type"] == 3, 4 + 16 * scene["shp_count"].astype(np.float64),
                          np.where(scene["shp_type"] == 0, 12, np.where(scene["shp_type"] == 1, 20, 16))).mean())
    alg = {
        # per-launch algorithmic bytes of each kernel (DESIGN.md "Roofline")
        "k_apply_translates": n_dyn * (12 + 4 + 16 + 16 + 16 + 16 + 16 + 16),
        "k_traverse<self>": n_dyn * (4 + 16 + 8) + (n_dyn - 1) * 64 + 0.8 * cand_per_frame * 8,
        "k_traverse<cross>": n_dyn * (4 + 16 + 8) + (n_sta - 1) * 64 + 0.2 * cand_per_frame * 8,
        # narrowphase: the polygon bin (polygon vs polygon|capsule|segment) holds ~45 % of the candidates of this mix
        "k_narrow_filter": 0.55 * cand_per_frame * (8 + 2 * (4 + 4 + 16 + 20)) + 0.2 * cand_per_frame * 16,
        "k_narrow_filter_pp": 0.45 * cand_per_frame * (8 + 2 * (4 + 4 + 16 + 4 + 16 * 5.5)) + 0.15 * cand_per_frame * 16,
        "k_narrow_manifold": 0.55 * pairs_per_frame / world * (16 + 2 * (4 + 4 + 16 + 20) + 84),
        "k_narrow_manifold_pp": 0.45 * pairs_per_frame / world * (16 + 2 * (4 + 4 + 16 + 4 + 16 * 5.5) + 84),
        "radix_sort(1 memset + 5 kernels)": n_dyn * (4 + 4 * 16),
        "k_fit": n_dyn * (4 + 16 + 8 + 4) + (n_dyn - 1) * (64 + 4 + 4),
        "k_karras_build": n_dyn * 4 + (n_dyn - 1) * (16 + 8),
        "k_morton": n_dyn * (4 + 16 + 8),
        "k_bounds": n_dyn * (4 + 16),
        "k_bin_count": cand_per_frame * (8 + 8),
        "k_bin_scatter": cand_per_frame * (8 + 8 + 8),
    }
    frame_alg = n_shapes * (20 + g_all + 16 + 16 + 8) + cand_per_frame * world * (16 + 2 * (20 + g_np)) + pairs_per_frame * 84
    dom_avg_ms = dom_ms / max(dom_count, 1)
    dom_bytes = float(alg.get(dom_name, 0.0))
    achieved = dom_bytes / (dom_avg_ms * 1e-3) / 1e9 if dom_avg_ms > 0 else 0.0
    kernels = {k: {"launches": c, "avg_ms": ms / max(c, 1), "share": ms / max(sum(v[1] for v in profile.values()), 1e-9),
                   "alg_gbs": (alg.get(k, 0.0) / (ms / max(c, 1) * 1e-3) / 1e9) if ms > 0 else 0.0}
               for k, (c, ms) in sorted(profile.items(), key=lambda kv: -kv[1][1])}
    # DRAM traffic per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)
    traffic, traffic_src = None, None
    tpath = os.path.join(REPO, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tpath):
        ncu_name = {"k_traverse<self>": "k_traverse<1>", "k_traverse<cross>": "k_traverse<0>"}.get(dom_name, dom_name)
        rec = json.load(open(tpath)).get(ncu_name)
        if rec:
            traffic = rec["dram"]
            traffic_src = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r01_ncu_full_top_kernels_v3.csv "
                           f"(cold-cache, serialised replay; {rec['launches_profiled']} launches)")
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "alg_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_avg_ms,
                "frame": {"alg_bytes": frame_alg, "achieved": frame_alg / (ms_per_step * 1e-3) / 1e9,
                          "frac": frame_alg / (ms_per_step * 1e-3) / 1e9 / peak},
                "note": "latency/divergence-bound path (SURVEY.md 8(d)): whole-frame algorithmic traffic is ~0.2 GB",
                "kernels": kernels}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {n_dyn} dynamic mixed circle/capsule/segment/polygon + {n_sta} static "
                               f"polygons, per-frame jitter U(-0.5,0.5)^2, LCG seed 4242 (SURVEY.md 8(d) cfg 3)",
                   "parallelism": ("single GPU" if world == 1 else
                                   f"weak: {world} x the workload ({n_total_dyn} dynamic + {n_total_sta} static in one world), one "
                                   f"equal-population x-slab per GPU + halo copies within {HALO_MARGIN} units, owner rule, no "
                                   f"data-path collective in the device-resident loop (moves are a function of the global id); the e2e "
                                   f"loop exchanges the boundary moves with one NCCL all-gather per frame "
                                   f"(rank 0 slab: {n_dyn} dynamic + {n_sta} static incl. halo)" if weak else
                                   f"strong: fixed workload mirrored on {world} GPUs, query leaves sharded by Morton-rank slab"),
                   "l2": "inputs larger than L2 (fresh 12 MB move batch per frame; ~255 MB shape state + ~100 MB "
                         "sort/tree buffers rewritten every frame vs 126 MB L2), no explicit flush",
                   "lbvh_rebuild_period": args.rebuild_period,
                   "broadphase": f"uniform grid, cell {args.grid_cell}" if args.grid_cell > 0 else "LBVH",
                   "pairs_per_frame": pairs_per_frame, "candidates_per_frame": cand_per_frame * world,
                   "stage_ms_per_step": {k: v / K for k, v in stage_ms.items()}},
        "e2e": {"value": e2e_pairs_all / e2e_elapsed, "unit": UNIT, "ms_per_step": 1e3 * e2e_elapsed / K,
                "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                "moves_host_ms": 1e3 * float(np.mean(move_times)),
                "gather_to_rank0_ms_per_step": 1e3 * float(np.mean(gather_times)) if gather_times else 0.0,
                "gather_note": "optional global view (all ranks' records gathered over NCCL + one download on rank 0); NOT "
                               "part of e2e.value: every rank returns the pairs of its own slab to its own host",
                "boundary_exchange_ms_per_step": 1e3 * float(np.mean(exchange_times)) if exchange_times else 0.0,
                "view_ms_per_step": 1e3 * float(np.mean(view_times)),
                "view_call": "Registry::getCollidingPairsView() (extension: span over the pinned result buffer)",
                "call": ("c2d::Registry<uint32_t>::getCollidingPairs() after moveEntity(id, Translation) per entity" if world == 1
                         else "per rank: boundary-move exchange (ExchangePlan: one NCCL all-gather of the 8-byte moves) + "
                              "c2d::Registry<uint32_t>::getCollidingPairs() of the rank's slab into host memory; time = max "
                              "over ranks, pairs = sum over ranks")},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
    }
    return out


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
    ap.add_argument("--workload", default="cfg3-clustered", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"],
                    help="multi-GPU mode: weak = the world grows with N (N x the workload, every rank owns one spatial slab "
                         "+ halo copies, no data-path collective); strong = the fixed workload mirrored on every rank, "
                         "query leaves sharded by Morton slab.  auto = weak for N > 1")
    ap.add_argument("--grid-cell", type=int, default=0,
                    help="> 0: use Registry<Grid>(cell) = the uniform-grid sort-and-sweep broadphase instead of the LBVH")
    ap.add_argument("--rebuild-period", type=int, default=4,
                    help="LBVH topology rebuilt every N-th frame, bottom-up refit in between (1 = every frame)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank, world, local = dist_env()

    if args.impl == "reference":
        if rank != 0:
            return 0
        weak = args.gpus > 1 and args.scaling in ("auto", "weak")
        scene, n_dyn, n_sta = build_scene(args.workload, args.gpus if weak else 1)
        r = run_reference(args, scene, n_dyn, as_baseline=False)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.workload}: {n_dyn} dynamic + {n_sta} static (same scene and seed as the own arm)",
                           "pairs_per_frame": r["pairs_per_frame"]},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores")},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), file=JSON_OUT, flush=True)
        return 0

    weak = world > 1 and args.scaling in ("auto", "weak")
    scene, n_dyn, n_sta = build_scene(args.workload, world if weak else 1)
    out = run_b200(args, scene, n_dyn, n_sta, rank, world, local)
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
