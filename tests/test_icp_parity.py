"""GPU parity: batched loop-closure ICP verification (PCL IterativeClosestPoint restatement) vs the CPU oracle.
Tolerance: the 17 correspondence sums are fp64 on both sides and the correspondences are exact 1-NN, so the
iteration counts agree and T agrees to ~1e-6; asserted <= 1e-4 rad / 1e-3 m like the registration loop."""
import numpy as np
import pytest

from lis_slam_b200 import synth
from oracle import orc

from common import local_map

pytestmark = pytest.mark.gpu


def _pair(seed, n_src=15000, noise=0.01, offset=(0.01, -0.02, 0.03, 0.4, -0.3, 0.1)):
    m = local_map()
    rng = np.random.default_rng(seed)
    Tt = synth.pose_to_T(np.array(offset))
    sel = rng.choice(len(m["surf"]), n_src, replace=False)
    src = m["surf"][sel].copy()
    src[:, :3] = ((src[:, :3].astype(np.float64) - Tt[:3, 3]) @ Tt[:3, :3] + rng.normal(0, noise, (n_src, 3))).astype(np.float32)
    return src, Tt


def _rot_err(Ta, Tb):
    # chordal distance ||Ra - Rb||_F / sqrt(2) ~ angle for small angles (arccos of a float32 trace is ill-conditioned near 0)
    dR = Ta[:3, :3].astype(np.float64) - Tb[:3, :3].astype(np.float64)
    return float(np.linalg.norm(dR) / np.sqrt(2)), float(np.linalg.norm(Ta[:3, 3] - Tb[:3, 3]))


def test_icp_batch_matches_oracle(engine):
    m = local_map()
    tgt = np.ascontiguousarray(m["surf"][::2])
    tid = engine.target_create(tgt)
    cases = [_pair(1), _pair(2, offset=(0.0, 0.0, 0.05, 0.8, 0.2, 0.0)), _pair(3, n_src=4000, offset=(0, 0, 0, 0, 0, 0))]
    res = engine.icp_verify_batch([(s, tid) for s, _ in cases])
    for (src, Tt), rg in zip(cases, res):
        To, ro = orc.icp(src, tgt)
        Tg = np.array(rg.T, np.float32).reshape(4, 4)
        assert rg.converged == ro.converged == 1
        assert rg.iters == ro.iters, (rg.iters, ro.iters)
        assert rg.n_corr_last == ro.n_corr_last
        er, et = _rot_err(To, Tg)
        assert er <= 1e-4 and et <= 1e-3, (er, et)
        assert abs(rg.fitness - ro.fitness) <= 1e-5 * max(1.0, ro.fitness)
        assert rg.fitness < 0.5                                   # historyKeyframeFitnessScore gate: a true loop
    engine.map_destroy(tid)


def test_icp_rejects_non_overlapping_and_handles_degenerate_inputs(engine):
    m = local_map()
    tgt = np.ascontiguousarray(m["surf"][::4])
    tid = engine.target_create(tgt)
    src, _ = _pair(4, n_src=3000)
    far = src.copy(); far[:, 0] += 500.0                            # no correspondences within 10 m
    two = src[:2].copy()                                            # < 3 correspondences
    res = engine.icp_verify_batch([(far, tid), (two, tid), (np.zeros((0, 4), np.float32), tid)])
    To, ro = orc.icp(far, tgt)
    assert res[0].converged == ro.converged == 0 and res[0].iters == ro.iters == 0
    assert abs(res[0].fitness - ro.fitness) <= 1e-4 * ro.fitness and res[0].fitness > 0.5   # unbounded-NN fitness: rejected
    To, ro = orc.icp(two, tgt)
    assert res[1].converged == ro.converged == 0
    assert res[2].converged == 0 and res[2].iters == 0
    engine.map_destroy(tid)
