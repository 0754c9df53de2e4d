"""GPU parity of the SECONDARY oracles of SURVEY.md section 8(c) item 5, bit for bit against the reference's own private
state over a multi-frame scripted scene:

* tight boxes    = ShapeInstance::mBoundingBox            (reference Registry.hpp:117, 191-231)
* leaf boxes     = mTrees[body].getNode(handle).mBoundingBox (DynamicBVH.hpp:352-365, 386-417: the fat-leaf rule)
* candidate set  = the five mTrees[..].findAllCollisions groups (DynamicBVH.hpp:931-1031, Registry.hpp:487-545)

The pair set stays exact for ANY conservative leaf box, so only these checks pin rows a2 / a3: a wrong fat-leaf rule would
show up as candidate inflation (lost time), never as a wrong result.  The reference side reads its private members through
the `#define private public` build of tests/cpp/runner.cpp (oracle/_ref); the B200 side reads c2d_gpu_read_boxes /
c2d_gpu_read_candidates through the C ABI.  tests/test_oracle_vs_reference.py runs the same bodies on the CPU with the C
oracle as the subject."""
import numpy as np
import pytest

from colli2de_b200 import scenes
from tests import harness as H

pytestmark = pytest.mark.gpu

ORACLE = "reference"
SUBJECT = "b200"


def _need_reference():
    if not H.available(ORACLE):
        pytest.skip("oracle/_ref (the compiled reference) is not available")


def _pair_keys(c: np.ndarray) -> np.ndarray:
    return np.sort((c[:, 0].astype(np.uint64) << np.uint64(32)) | c[:, 1].astype(np.uint64))


def check_state(ref, sub, label=""):
    """tight + leaf boxes of every shape slot and the candidate set: bit-identical / identical."""
    n = ref.shape_slots()
    assert sub.shape_slots() == n, label
    rt, rf, ra = ref.read_boxes(0, n)
    st, sf, sa = sub.read_boxes(0, n)
    assert np.array_equal(ra, sa), f"{label}: alive flags differ"
    live = ra.astype(bool)
    bad_t = np.nonzero((rt.view(np.uint32) != st.view(np.uint32)).any(axis=1) & live)[0]
    assert len(bad_t) == 0, f"{label}: {len(bad_t)} tight boxes differ, first shape {bad_t[:5]}: {rt[bad_t[:2]]} vs {st[bad_t[:2]]}"
    bad_f = np.nonzero((rf.view(np.uint32) != sf.view(np.uint32)).any(axis=1) & live)[0]
    assert len(bad_f) == 0, f"{label}: {len(bad_f)} leaf boxes differ, first shape {bad_f[:5]}: {rf[bad_f[:2]]} vs {sf[bad_f[:2]]}"
    rc, sc = ref.candidates(), sub.candidates()
    rk, sk = _pair_keys(rc), _pair_keys(sc)
    assert len(np.unique(rk)) == len(rk), f"{label}: the reference's candidate list has duplicates"
    assert len(np.unique(sk)) == len(sk), f"{label}: candidate list has duplicates"
    assert len(rk) == len(sk) and np.array_equal(rk, sk), (
        f"{label}: candidate sets differ: reference {len(rk)}, subject {len(sk)}, "
        f"missing {len(np.setdiff1d(rk, sk))}, extra {len(np.setdiff1d(sk, rk))}")
    return int(live.sum()), len(rk)


def run_scripted_scene(n_dyn=6000, n_static=1500, side=220.0, frames=6):
    """The D12 probe of SURVEY.md appendix B item 4, widened: alternating translation-only and translation+rotation moves,
    bullets with large steps, teleports, removals with ShapeId recycling, re-insertions."""
    _need_reference()
    sc = scenes.mixed_scene(n_dyn, n_static, side, seed=4242)
    sc["ent_body"][::31] = 2
    ref, sub = H.Backend(ORACLE), H.Backend(SUBJECT)
    for b in (ref, sub):
        scenes.populate(b, sc)
    rng = scenes.LCG(77)
    n = len(sc["ent_id"])
    ids = np.arange(n, dtype=np.uint32)
    body = sc["ent_body"].copy()
    alive = np.ones(n, bool)
    totals = []
    totals.append(check_state(ref, sub, "after insert"))
    for frame in range(frames):
        mov = ids[(body != 0) & alive]
        if frame % 2 == 0:
            d = scenes.jitter(rng, len(mov), 0.8)
            for b in (ref, sub):
                b.move_translate(mov, d)
        else:
            nb = ids[(body == 1) & alive]
            d3 = np.zeros((len(nb), 3), np.float32)
            d3[:, :2] = scenes.jitter(rng, len(nb), 0.5)
            d3[:, 2] = rng.next_float(-0.4, 0.4, len(nb))
            for b in (ref, sub):
                b.move_transform(nb, d3)
        bul = ids[(body == 2) & alive]
        big = scenes.jitter(rng, len(bul), 12.0)
        for b in (ref, sub):
            b.move_translate(bul, big)
        # the same entity moved twice before one flush: the leaf rule is history dependent
        twice = mov[frame::53]
        d2 = scenes.jitter(rng, len(twice), 1.5)
        for b in (ref, sub):
            b.move_translate(twice, d2)
        tel = ids[(body == 1) & alive][frame::71]
        t3 = np.zeros((len(tel), 3), np.float32)
        t3[:, :2] = rng.next_vec2(0.0, side, len(tel))
        t3[:, 2] = rng.next_float(0.0, 6.28, len(tel))
        for b in (ref, sub):
            b.teleport_transform(tel, t3)
        if frame == 2:
            gone = ids[alive][3::17]
            for b in (ref, sub):
                b.remove_entities(gone)
            alive[gone] = False
        if frame == 3:
            # new entities reuse the freed ShapeIds LIFO (Registry.hpp:283-289)
            fresh = np.arange(n, n + 200, dtype=np.uint32)
            xf3 = np.zeros((200, 3), np.float32)
            xf3[:, :2] = rng.next_vec2(0.0, side, 200)
            g = np.zeros((200, 16), np.float32)
            g[:, 2] = 1.25
            for b in (ref, sub):
                b.create_entities(fresh, np.full(200, H.DYNAMIC, np.uint8), xf3)
                b.add_shapes(fresh, np.zeros(200, np.uint8), g)
            ids = np.concatenate([ids, fresh])
            body = np.concatenate([body, np.full(200, 1, np.uint8)])
            alive = np.concatenate([alive, np.ones(200, bool)])
            n += 200
        totals.append(check_state(ref, sub, f"frame {frame}"))
    return totals


def test_boxes_and_candidates_scripted_scene():
    totals = run_scripted_scene()
    assert all(c > 10000 for _, c in totals), totals


def test_boxes_and_candidates_cfg1b_frames():
    """BASELINE cfg 1b (10 k dynamic circles + 1 k static polygons): four frames of jitter."""
    _need_reference()
    sc = scenes.baseline_cfg1()
    ref, sub = H.Backend(ORACLE), H.Backend(SUBJECT)
    for b in (ref, sub):
        scenes.populate(b, sc)
    rng = scenes.LCG(5)
    dyn = np.nonzero(sc["ent_body"] == 1)[0].astype(np.uint32)
    for frame in range(4):
        live, cands = check_state(ref, sub, f"frame {frame}")
        assert live == 11000 and cands > 10000
        d = scenes.jitter(rng, len(dyn), 0.5)
        for b in (ref, sub):
            b.move_translate(dyn, d)


def test_boxes_and_candidates_100k_tree_pipeline():
    """cfg 2 at its stated size (100 k mixed dynamic + 10 k static): the LBVH pipeline's candidates against the
    reference's DynamicBVH candidates over three frames."""
    _need_reference()
    sc = scenes.mixed_scene(100000, 10000, 1600.0)
    ref, sub = H.Backend(ORACLE), H.Backend(SUBJECT)
    for b in (ref, sub):
        scenes.populate(b, sc)
    rng = scenes.LCG(6)
    dyn = np.nonzero(sc["ent_body"] == 1)[0].astype(np.uint32)
    for frame in range(3):
        d = scenes.jitter(rng, len(dyn), 0.5)
        for b in (ref, sub):
            b.move_translate(dyn, d)
        live, cands = check_state(ref, sub, f"frame {frame}")
        assert live == 110000 and cands > 50000
