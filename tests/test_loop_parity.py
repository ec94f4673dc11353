"""GPU parity: the device loop detector (lisreg_loop_detect: project + globalICP + per-candidate re-description +
scoring + bookkeeping, epscGeneration.cpp:84-120, :258-401, :663-992) against the CPU restatement on the same
there-and-back keyframe sequence.  BIT-EXACT: frame ids, candidate counts, scores (integer SADs of u8 descriptors)
and the 4x4 transforms are equal - the 17 fp64 sums of the 2-D ICP are accumulated in source-index order on both
sides (loop.cuh k_loop_align / oracle orc_icp.cpp), so the fitted transform and every byte binned after it agree."""
import numpy as np
import pytest

from lis_slam_b200 import engine as E
from oracle import orc

from common import loop_keyframes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("flags", [dict(use_fepsc=True), dict(use_epsc=True, use_sepsc=True, use_fepsc=True, use_pose=True)])
def test_loop_detect_matches_oracle(engine, flags):
    kfs = loop_keyframes()
    lut = orc.using_map_lut()
    det_o = orc.LoopDetector(lut=lut, **flags)
    det_g = engine.loop_create(lut, **flags)
    n_match = 0
    for k, (corner, surf, sem, lab, odom) in enumerate(kfs):
        co, no, mo = det_o.detect(corner, surf, sem, lab, odom)
        cg, ng, mg = engine.loop_detect(det_g, corner, surf, sem, lab, odom)
        assert (cg, ng) == (co, no) == (k, no)
        assert [(a, b) for a, b, _, _ in mg] == [(a, b) for a, b, _, _ in mo], (k, mg, mo)
        for (kind, mid, so, To), (_, _, sg, Tg) in zip(mo, mg):
            assert so == sg, (k, kind, so, sg)
            assert np.array_equal(To, Tg), (k, kind, To, Tg)
            n_match += 1
    assert n_match >= 8
    det_o.close()
    engine.loop_destroy(det_g)


def test_loop_detect_edge_cases(engine):
    """Empty clouds and a jumpy trajectory: the gate (which measures the distance to the PREVIOUS keyframe, since
    posArr.back() is pushed after the loop, epscGeneration.cpp:740 / :899), the ids and the matches follow the oracle."""
    lut = orc.using_map_lut()
    det = engine.loop_create(lut, use_fepsc=True, use_pose=True)
    od = orc.LoopDetector(lut=lut, use_fepsc=True, use_pose=True)
    empty = np.zeros((0, 4), np.float32); nolab = np.zeros(0, np.uint16)
    rng = np.random.default_rng(3)
    few = np.zeros((40, 4), np.float32); few[:, 0] = rng.uniform(5, 20, 40); few[:, 1] = rng.uniform(-5, 5, 40)
    fewlab = np.full(40, 18, np.uint16)
    xs = [0.0, 30.0, 60.0, 0.0, 0.0, 30.2, 59.9, 0.1]
    total_cand = 0
    for k, x in enumerate(xs):
        T = np.eye(4, dtype=np.float32); T[0, 3] = x
        clouds = (empty, empty, empty, nolab) if k % 2 == 0 else (few[:10], few, few, fewlab)
        cg, ng, mg = engine.loop_detect(det, *clouds, T)
        co, no, mo = od.detect(*clouds, T)
        assert (cg, ng) == (co, no) == (k, no)
        assert [(a, b) for a, b, _, _ in mg] == [(a, b) for a, b, _, _ in mo], (k, mg, mo)
        for (_, _, so, To), (_, _, sg, Tg) in zip(mo, mg):
            assert so == sg and np.array_equal(To, Tg)
        total_cand += ng
    assert total_cand > 0
    od.close()
    engine.loop_destroy(det)


def test_loop_history_grows_past_its_first_allocation(engine):
    """300 keyframes (first device allocation = 256): the history survives the re-allocation - a revisit of keyframe 0
    after the growth still matches it with the same score as the oracle."""
    lut = orc.using_map_lut()
    det = engine.loop_create(lut, use_fepsc=True)
    od = orc.LoopDetector(lut=lut, use_fepsc=True)
    rng = np.random.default_rng(9)
    few = np.zeros((60, 4), np.float32); few[:, 0] = rng.uniform(5, 40, 60); few[:, 1] = rng.uniform(-20, 20, 60)
    lab = np.where(np.arange(60) % 2 == 0, 13, 18).astype(np.uint16)
    last = None
    for k in range(300):
        T = np.eye(4, dtype=np.float32); T[0, 3] = 0.5 * k
        pts = few.copy(); pts[:, 1] += 0.01 * (k % 7)
        a = engine.loop_detect(det, pts[:20], pts, pts, lab, T)
        b = od.detect(pts[:20], pts, pts, lab, T)
        assert a[:2] == b[:2]
    for rep in range(2):                               # jump back to the start: the second call sees the start gated
        T = np.eye(4, dtype=np.float32); T[0, 3] = 0.01 * rep
        cg, ng, mg = engine.loop_detect(det, few[:20], few, few, lab, T)
        co, no, mo = od.detect(few[:20], few, few, lab, T)
        assert (cg, ng) == (co, no)
        assert [(x[0], x[1]) for x in mg] == [(x[0], x[1]) for x in mo]
        for (_, _, so, To), (_, _, sg, Tg) in zip(mo, mg):
            assert so == sg and np.array_equal(To, Tg)
        last = (ng, mg)
    assert last[0] > 0
    od.close(); engine.loop_destroy(det)
