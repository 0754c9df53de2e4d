"""Multi-GPU group on REAL ranks (one process per GPU over NCCL): run under torchrun, e.g.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py

Every rank holds a c2d::Registry of the same scene joined to one group (Registry::joinGroup).  Frame by frame the ranks
move the entities they drive (move exchange: ncclAllGather of the 12-byte records), call the collective
getCollidingPairs() in each result mode, and rank 0 compares the union with the UNMODIFIED reference (oracle/_ref):
exact tuple set, manifolds, ambiguous segment pairs classified (tests/parity.py).  Structural changes in the middle are
replicated on every rank.  Also checked: the host mirror of entities moved by OTHER ranks (getEntityTransform after an
exchange) is bit-identical to the reference's transforms.  Prints MULTIGPU_OK on rank 0.

TEST INFRASTRUCTURE (tests/test_gpu_multigpu.py launches it when the box has >= 2 GPUs)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from colli2de_b200 import capi, scenes  # noqa: E402
from tests import harness as H, parity as P  # noqa: E402


def gather_numpy(arr: np.ndarray):
    """all ranks' structured arrays concatenated on every rank (python-side union for the LOCAL result mode)"""
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, arr.tobytes())
    return np.concatenate([np.frombuffer(p, dtype=arr.dtype) for p in parts])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_dyn, n_sta = int(os.environ.get("C2D_CHECK_DYN", "120000")), int(os.environ.get("C2D_CHECK_STA", "12000"))
    sc = scenes.mixed_scene(n_dyn, n_sta, 1700.0 * (n_dyn / 120000.0) ** 0.5, seed=23, clustered=True)
    sc["ent_body"][:n_dyn:41] = 2  # bullets among the dynamic entities
    which = "reference" if H.available("reference") else "oracle"
    b = H.Backend("b200")
    scenes.populate(b, sc)
    b.group_join_torch()
    ctx = b.native_context()
    capi.set_slab_mode(ctx, 1000)
    ref, idx = None, None
    if rank == 0:
        ref = H.Backend(which)
        scenes.populate(ref, sc)
        idx = P.SceneIndex(sc)
    movable = np.nonzero(sc["ent_body"] != 0)[0].astype(np.uint32)
    alive = np.ones(len(movable), bool)
    rng = scenes.LCG(99)
    modes = [(capi.RESULTS_ROOT, True), (capi.RESULTS_ALL, True), (capi.RESULTS_LOCAL, True), (capi.RESULTS_ALL, False),
             (capi.RESULTS_ROOT, True), (capi.RESULTS_LOCAL, False)]
    totals = dict(expected=0, ambiguous=0, not_bit_identical=0)
    for f, (mode, exchange) in enumerate(modes):
        d = scenes.jitter(rng, len(movable), 0.6)
        if f == 3:
            # structural change, replicated on every rank: entities leave, shapes are switched off
            gone = movable[200:260]
            off = np.arange(5000, 5040, dtype=np.uint32)
            for x in [b] + ([ref] if ref else []):
                x.remove_entities(gone)
                x.set_active(off, np.zeros(len(off), np.uint8))
            alive[200:260] = False
        ids, dd = movable[alive], d[alive]
        b.group_options(results=mode, exchange=exchange)
        if exchange:  # this rank drives one block of the entities
            lo, hi = (len(ids) * rank) // world, (len(ids) * (rank + 1)) // world
            b.move_translate(ids[lo:hi], dd[lo:hi])
        else:
            b.move_translate(ids, dd)
        got = b.get_colliding_pairs()
        st = capi.comm_stats(ctx)
        if mode == capi.RESULTS_LOCAL:
            got = gather_numpy(got)
        elif mode == capi.RESULTS_ROOT and rank != 0:
            assert len(got) == st["local_pairs"]
        counts = [None] * world
        dist.all_gather_object(counts, int(st["local_pairs"]))
        assert min(counts) > 0, counts
        if exchange:
            assert st["exchange_bytes"] == 12 * (len(ids) - (hi - lo)), (st, len(ids), hi - lo)
        assert st["slab_owned"] > 0 and st["slab_leaves"] >= st["slab_owned"]
        if rank == 0:
            ref.move_translate(ids, dd)
            res = P.compare_pairs(which, ref, idx, ref.get_colliding_pairs(), got, max_ambiguous=20)
            assert res["expected"] > 20000 and res["not_bit_identical"] == 0, res
            assert sum(counts) == res["got"]
            for k in totals:
                totals[k] += res[k]
        if exchange and f in (1, 4):
            # host mirror of entities other ranks moved: read back lazily from the device, equal to the reference's
            probe = ids[:: max(1, len(ids) // 500)]
            mine = b.get_transforms(probe)
            want = [mine.tobytes() if rank != 0 else ref.get_transforms(probe).tobytes()]
            dist.broadcast_object_list(want, src=0)
            assert mine.tobytes() == want[0], "host transforms differ from the reference after a move exchange"
    # one-step exchange (blocks sized from the previous frame) and its overflow path: two ordinary frames, then rank 0
    # suddenly drives EVERY entity (its count exceeds the agreed block: nothing is applied, the frame is redone exactly),
    # then a frame in which only the last rank moves anything
    b.group_options(results=capi.RESULTS_ROOT, exchange=True)
    ids = movable[alive]
    for step in range(5):
        dd = scenes.jitter(rng, len(ids), 0.5)
        lo, hi = (len(ids) * rank) // world, (len(ids) * (rank + 1)) // world
        if step == 2:
            lo, hi = (0, len(ids)) if rank == 0 else (0, 0)
        if step == 3:
            lo, hi = (0, len(ids)) if rank == world - 1 else (0, 0)
        b.move_translate(ids[lo:hi], dd[lo:hi])
        print(f"[rank {rank}] one-step exchange frame {step}: {hi - lo} moves logged", file=sys.stderr, flush=True)
        got = b.get_colliding_pairs()
        print(f"[rank {rank}] frame {step} returned {len(got)} records, retries {int(b.stats()['retries'])}", file=sys.stderr, flush=True)
        st = capi.comm_stats(ctx)
        assert st["exchange_bytes"] == 12 * (len(ids) - (hi - lo)), (step, st)
        # the host mirror of the entities other ranks moved is refreshed by the first call that reads a transform; that has
        # to happen before this rank logs its next moves (the runner reads bullets' transforms while it moves them)
        b.get_transforms(ids[:1])
        retries = [None] * world
        dist.all_gather_object(retries, int(b.stats()["retries"]))
        assert all(r >= 1 for r in retries) if step == 2 else True, (step, retries)  # every rank redid the overflowing frame
        if rank == 0:
            ref.move_translate(ids, dd)
            res = P.compare_pairs(which, ref, idx, ref.get_colliding_pairs(), got, max_ambiguous=20)
            assert res["not_bit_identical"] == 0, res
            totals["expected"] += res["expected"]
    b.group_options(results=capi.RESULTS_LOCAL, exchange=False)
    # batched rayCast with ray sharding: every rank is handed the same batch and casts its contiguous share; the per-ray
    # lists of the ranks concatenate to the reference's rayCast results
    ray4, inf = scenes.rays(1200, 1700.0 * (n_dyn / 120000.0) ** 0.5, seed=31, finite_len=60.0)
    b.group_ray_sharding(True)
    off, hits = b.raycast(ray4, inf)
    counts = (off[1:] - off[:-1]).astype(np.int64)
    lo, hi = (len(ray4) * rank) // world, (len(ray4) * (rank + 1)) // world
    assert counts[:lo].sum() == 0 and counts[hi:].sum() == 0 and counts[lo:hi].sum() == len(hits) > 0
    all_counts, all_hits = [None] * world, [None] * world
    dist.all_gather_object(all_counts, counts[lo:hi].tobytes())
    dist.all_gather_object(all_hits, hits.tobytes())
    if rank == 0:
        got_counts = np.concatenate([np.frombuffer(c, np.int64) for c in all_counts])
        got_hits = np.concatenate([np.frombuffer(h, hits.dtype) for h in all_hits])
        roff, rhits = ref.raycast(ray4, inf)
        got_off = np.concatenate([[0], np.cumsum(got_counts)])
        same = sum(1 for i in range(len(ray4)) if rhits[int(roff[i]):int(roff[i + 1])].tobytes() ==
                   got_hits[int(got_off[i]):int(got_off[i + 1])].tobytes())
        assert same >= len(ray4) - 4, f"only {same} of {len(ray4)} rays agree with the {which}"  # exact (entry, exit) ties (D10)
        totals["rays_equal"] = same
        totals["ray_hits"] = int(len(rhits))
    b.group_ray_sharding(False)
    b.group_leave()
    # out of the group the registry is a plain single-GPU registry again
    single = b.get_colliding_pairs()
    if rank == 0:
        res = P.compare_pairs(which, ref, idx, ref.get_colliding_pairs(), single, max_ambiguous=20)
        print("MULTIGPU_OK", world, totals, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
