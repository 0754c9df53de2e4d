"""CPU, world_size 2 over gloo: the multi-GPU plumbing around the path — Morton-rank slab bounds
(c2d_gpu_set_shard uses the same arithmetic), the owner rule (a pair belongs to the rank that owns its query
leaf) and the variable-size gather of 84-byte pair records to rank 0.  Each rank runs the C oracle on the full
scene, keeps the pairs its slab owns, and rank 0 must end up with exactly the full set, no duplicates."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from colli2de_b200 import scenes, sharding
from tests import harness as H, parity as P


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.mixed_scene(3000, 300, 260.0, seed=5)
    b = H.Backend("oracle")
    scenes.populate(b, sc)
    pairs = b.get_colliding_pairs()
    # query-leaf order stand-in: shapes sorted by id (the GPU uses Morton rank; the rule is the same)
    n_query = 3300
    lo, hi = sharding.slab_bounds(n_query, rank, world)
    mine = pairs[(pairs["shapeA"] >= lo) & (pairs["shapeA"] < hi)]
    gathered = sharding.gather_records(mine, dst=0)
    total = sharding.allreduce_sum(len(mine))
    if rank == 0:
        np.save(out_path, gathered)
        assert total == len(pairs)
        assert len(gathered) == len(pairs)
        assert P.sort_pairs(gathered).tobytes() == P.sort_pairs(pairs).tobytes()
    dist.destroy_process_group()


@pytest.mark.skipif(not H.available("oracle"), reason="needs oracle/libc2d_oracle.so")
def test_slab_sharding_and_gather_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path / "gathered.npy")), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "gathered.npy")


def test_slab_bounds_partition():
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            edges = [sharding.slab_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _ghost_worker(rank, world, port, out_dir):
    """Weak-scaling decomposition: each rank holds ONLY its slab + halo copies; the owner rule must make the union of
    the ranks' results equal to the single-registry result, each pair once."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.mixed_scene(4000, 400, 300.0, seed=15, clustered=True)
    sc["ent_body"][::30] = 2
    x = sc["ent_xf3"][:, 0]
    owned, ghost = sharding.slab_selection(x, rank, world, margin=40.0)
    local = np.nonzero(owned | ghost)[0]
    sub = {k: v[local] for k, v in sc.items()}
    b = H.Backend("oracle")
    ids = scenes.populate(b, sub)  # entity ids stay global, shape ids are local
    b.set_ghost(ids[ghost[local]], np.ones(int(ghost[local].sum()), np.uint8))
    rng = scenes.LCG(3)
    for frame in range(3):
        d = scenes.jitter(rng, len(sc["ent_id"]), 0.6)  # every rank can generate any shape's move (replicated input)
        movable = sc["ent_body"][local] != 0
        b.move_translate(sc["ent_id"][local][movable], d[local][movable])
    mine = b.get_colliding_pairs()
    keys = np.stack([np.minimum(mine["entityA"], mine["entityB"]), np.maximum(mine["entityA"], mine["entityB"])], 1)
    gathered = sharding.gather_records(np.ascontiguousarray(keys).view([("a", "<u4"), ("b", "<u4")]).ravel(), dst=0)
    if rank == 0:
        full = H.Backend("oracle")
        scenes.populate(full, sc)
        rng = scenes.LCG(3)
        for frame in range(3):
            d = scenes.jitter(rng, len(sc["ent_id"]), 0.6)
            movable = sc["ent_body"] != 0
            full.move_translate(sc["ent_id"][movable], d[movable])
        fp = full.get_colliding_pairs()
        fk = np.stack([np.minimum(fp["entityA"], fp["entityB"]), np.maximum(fp["entityA"], fp["entityB"])], 1)
        want = np.sort(fk[:, 0].astype(np.uint64) << np.uint64(32) | fk[:, 1])
        got = np.sort(gathered["a"].astype(np.uint64) << np.uint64(32) | gathered["b"])
        assert len(got) == len(np.unique(got)), "a pair was reported by two ranks"
        assert np.array_equal(want, got), (len(want), len(got))
        assert len(want) > 1000
        open(os.path.join(out_dir, "ok"), "w").write(str(len(want)))
    dist.destroy_process_group()


@pytest.mark.skipif(not H.available("oracle"), reason="needs oracle/libc2d_oracle.so")
@pytest.mark.parametrize("world", [2, 3])
def test_slabs_with_ghosts_and_owner_rule(tmp_path, world):
    port = _free_port()
    mp.spawn(_ghost_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(tmp_path / "ok")


def _exchange_worker(rank, world, port, out_dir, use_plan=False):
    """Like _ghost_worker, but a rank only KNOWS the moves of the shapes it owns: the halo copies receive theirs through
    sharding.exchange_boundary_moves (the data-path collective of the decomposition)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenes.mixed_scene(4000, 400, 300.0, seed=25, clustered=False)
    x = sc["ent_xf3"][:, 0]
    margin = 40.0
    owned, ghost = sharding.slab_selection(x, rank, world, margin)
    exports = sharding.slab_exports(x, rank, world, margin)
    local = np.nonzero(owned | ghost)[0]
    b = H.Backend("oracle")
    ids = scenes.populate(b, {k: v[local] for k, v in sc.items()})
    b.set_ghost(ids[ghost[local]], np.ones(int(ghost[local].sum()), np.uint8))
    movable = sc["ent_body"] != 0
    ghost_ids = np.sort(sc["ent_id"][ghost & movable])
    plan = sharding.ExchangePlan(sc["ent_id"][exports & movable], ghost_ids) if use_plan else None
    if plan is not None:
        assert plan.found.all(), "a halo copy has no exporting owner"
    for frame in range(3):
        # moves of the OWNED shapes only (a private stream per rank would do; here a function of id and frame)
        mine = sc["ent_id"][owned & movable]
        r = np.random.default_rng(1000 * frame + 7)
        d_all = r.uniform(-0.6, 0.6, (len(sc["ent_id"]), 2)).astype(np.float32)  # stands in for "what the owner decided"
        d_mine = d_all[mine]
        exp = sc["ent_id"][exports & movable]
        if plan is not None:
            # persistent plan: only the 8-byte moves travel, the ghosts' rows come back in ghost_ids order
            d_ghost = plan.exchange(d_all[exp])
            assert np.array_equal(d_ghost, d_all[ghost_ids])
            b.move_translate(np.concatenate([mine, ghost_ids]), np.concatenate([d_mine, d_ghost]))
            continue
        all_ids, all_d = sharding.exchange_boundary_moves(exp, d_all[exp])
        sel = np.isin(all_ids, ghost_ids)
        assert set(all_ids[sel].tolist()) == set(ghost_ids.tolist()), "a halo copy did not receive its move"
        b.move_translate(np.concatenate([mine, all_ids[sel]]), np.concatenate([d_mine, all_d[sel]]))
    p = b.get_colliding_pairs()
    keys = np.stack([np.minimum(p["entityA"], p["entityB"]), np.maximum(p["entityA"], p["entityB"])], 1)
    gathered = sharding.gather_records(np.ascontiguousarray(keys).view([("a", "<u4"), ("b", "<u4")]).ravel(), dst=0)
    if rank == 0:
        full = H.Backend("oracle")
        scenes.populate(full, sc)
        for frame in range(3):
            r = np.random.default_rng(1000 * frame + 7)
            d_all = r.uniform(-0.6, 0.6, (len(sc["ent_id"]), 2)).astype(np.float32)
            full.move_translate(sc["ent_id"][movable], d_all[movable])
        fp = full.get_colliding_pairs()
        fk = np.stack([np.minimum(fp["entityA"], fp["entityB"]), np.maximum(fp["entityA"], fp["entityB"])], 1)
        want = np.sort(fk[:, 0].astype(np.uint64) << np.uint64(32) | fk[:, 1])
        got = np.sort(gathered["a"].astype(np.uint64) << np.uint64(32) | gathered["b"])
        assert np.array_equal(want, got), (len(want), len(got))
        open(os.path.join(out_dir, "ok"), "w").write(str(len(want)))
    dist.destroy_process_group()


@pytest.mark.skipif(not H.available("oracle"), reason="needs oracle/libc2d_oracle.so")
def test_boundary_move_exchange_world2(tmp_path):
    port = _free_port()
    mp.spawn(_exchange_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok")


@pytest.mark.skipif(not H.available("oracle"), reason="needs oracle/libc2d_oracle.so")
def test_exchange_plan_world3(tmp_path):
    """sharding.ExchangePlan: the halo-exchange plan built once, then one fixed-shape all-gather per frame."""
    port = _free_port()
    mp.spawn(_exchange_worker, args=(3, port, str(tmp_path), True), nprocs=3, join=True)
    assert os.path.exists(tmp_path / "ok")
