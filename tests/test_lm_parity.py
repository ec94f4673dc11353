"""GPU parity tests proper: the CUDA path (through the C-ABI) against the CPU oracle on the same
seeded inputs.  Tolerances: kNN bit-exact; per-iteration selection counts equal (<= 0.1 % slack,
SURVEY.md §8c); final pose <= 1e-4 rad / 1e-3 m (BASELINE.json north_star) — in practice ~1e-6."""
import numpy as np
import pytest

from lis_slam_b200 import engine as E
from lis_slam_b200 import synth
from oracle import orc

from common import lattice_map, lattice_queries, local_map, reg_case, scene

pytestmark = pytest.mark.gpu

ROT_TOL, TRANS_TOL = 1e-4, 1e-3


def _orc_params(variant, **kw):
    return orc.lm_params(variant, **kw)


def _check_against_oracle(engine, mid, f, guess, variant, m, early_exit=1, max_iters=None, labels=False):
    kw = {"early_exit": early_exit}
    if max_iters:
        kw["max_iters"] = max_iters
    po = _orc_params(variant, **kw)
    pg = E.lm_params(variant, **kw)
    cl = f["corner_label"] if labels else None
    sl = f["surf_label"] if labels else None
    pose_o, res_o, log_o = orc.scan2map(f["corner"], f["surf"], m["corner"], m["surf"], guess, po, clabel=cl, slabel=sl)
    pose_g, res_g, log_g = engine.scan2map(mid, f["corner"], f["surf"], guess, pg, clabel=cl, slabel=sl, log=True)
    assert res_g.status == res_o.status
    assert res_g.iters == res_o.iters, (res_g.iters, res_o.iters)
    assert res_g.converged == res_o.converged
    assert res_g.is_degenerate == res_o.is_degenerate
    for lo, lg in zip(log_o, log_g):
        assert abs(lo.n_sel - lg.n_sel) <= max(1, int(0.001 * lo.n_sel)), (lo.n_sel, lg.n_sel)
        A_o, A_g = np.array(lo.AtA), np.array(lg.AtA)
        assert np.abs(A_o - A_g).max() <= 2e-4 * np.abs(A_o).max()
        assert np.abs(np.array(lo.pose) - np.array(lg.pose))[:3].max() <= ROT_TOL
        assert np.abs(np.array(lo.pose) - np.array(lg.pose))[3:].max() <= TRANS_TOL
    er, et = synth.pose_error(pose_o, pose_g)
    assert er <= ROT_TOL and et <= TRANS_TOL, (er, et)
    return pose_o, pose_g, res_o, res_g, log_o, log_g


@pytest.fixture(scope="module")
def gpu_map(engine):
    m = local_map()
    mid = engine.map_create(m["corner"], m["surf"], gate_hint=2.0)
    yield mid, m
    engine.map_destroy(mid)


@pytest.fixture(scope="module")
def gpu_map_a(engine):
    m = local_map()
    mid = engine.map_create(m["corner"], m["surf"], gate_hint=1.0)
    yield mid, m
    engine.map_destroy(mid)


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("gate", [1.0, 2.0])
def test_knn5_bit_exact(engine, gpu_map, which, gate):
    mid, m = gpu_map
    cloud = m["corner"] if which == 0 else m["surf"]
    f, truth, _ = reg_case(0)
    T = synth.pose_to_T(truth)
    src = f["corner"] if which == 0 else f["surf"]
    q = np.zeros((len(src), 4), np.float32)
    q[:, :3] = (src[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    # add far-away and boundary queries
    q[:50, :3] += 500.0
    q[50:100, 2] += 3.0
    idx_o, sqd_o = orc.knn(cloud, q, 5)
    idx_g, sqd_g = engine.knn5(mid, which, q, gate)
    inside = sqd_o < gate
    assert np.array_equal(np.where(inside, idx_o, -1), idx_g)
    assert np.array_equal(sqd_o[inside], sqd_g[inside])
    assert inside[:, 4].sum() > 0.5 * len(q)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_variant_a_matches_oracle(engine, gpu_map_a, seed):
    mid, m = gpu_map_a
    f, truth, guess = reg_case(seed)
    pose_o, pose_g, res_o, res_g, *_ = _check_against_oracle(engine, mid, f, guess, "A", m)
    # and both recover the ground truth to the noise floor
    er, et = synth.pose_error(truth, pose_g)
    assert er < 2e-3 and et < 2e-2


def test_variant_a_fixed_10_iterations(engine, gpu_map_a):
    mid, m = gpu_map_a
    f, truth, guess = reg_case(3)
    _, _, res_o, res_g, *_ = _check_against_oracle(engine, mid, f, guess, "A", m, early_exit=0, max_iters=10)
    assert res_g.iters == 10


@pytest.mark.parametrize("variant", ["B", "C"])
def test_variants_b_c_label_weighted(engine, gpu_map, variant):
    mid, m = gpu_map
    f, truth, guess = reg_case(4)
    _check_against_oracle(engine, mid, f, guess, variant, m, labels=True)


def test_batch_matches_single_and_oracle(engine, gpu_map_a):
    mid, m = gpu_map_a
    cases = [reg_case(s, n_corner=1500 + 200 * s, n_surf=5000 + 300 * s) for s in range(40)]
    regs = [(mid, f["corner"], f["surf"], None, None) for f, _, _ in cases]
    guesses = [g for _, _, g in cases]
    p = E.lm_params("A")
    poses, res, _ = engine.scan2map_batch(regs, guesses, p)
    for s in (0, 7, 39):
        f, truth, guess = cases[s]
        pose_o, res_o, _ = orc.scan2map(f["corner"], f["surf"], m["corner"], m["surf"], guess, orc.lm_params("A"), log=False)
        er, et = synth.pose_error(pose_o, poses[s])
        assert er <= ROT_TOL and et <= TRANS_TOL
        assert res[s].iters == res_o.iters
        p1, r1, _ = engine.scan2map(mid, f["corner"], f["surf"], guess, E.lm_params("A"))
        assert np.array_equal(p1, poses[s])   # deterministic reduction: batch == single, bit for bit


def test_not_enough_features_leaves_pose_untouched(engine, gpu_map_a):
    mid, m = gpu_map_a
    f, truth, guess = reg_case(5)
    p = E.lm_params("A")
    pose, res, _ = engine.scan2map(mid, f["corner"], f["surf"][:100], guess, p)   # ns > 100 fails
    assert res.status == E.NOT_ENOUGH_FEATURES and res.iters == 0
    assert np.array_equal(pose, np.asarray(guess, np.float32))
    po, ro, _ = orc.scan2map(f["corner"], f["surf"][:100], m["corner"], m["surf"], guess, orc.lm_params("A"))
    assert ro.status == 1


def test_few_correspondences_is_noop(engine, gpu_map_a):
    """< 50 matches => every iteration is a silent no-op (odomEstimationNode.cpp:870-872)."""
    mid, m = gpu_map_a
    f, truth, guess = reg_case(6)
    far = np.array(guess, np.float32); far[3] += 500.0
    p = E.lm_params("A")
    pose, res, _ = engine.scan2map(mid, f["corner"], f["surf"], far, p)
    po, ro, _ = orc.scan2map(f["corner"], f["surf"], m["corner"], m["surf"], far, orc.lm_params("A"))
    assert res.status == E.FEW_CORRESPONDENCES == ro.status
    assert res.iters == ro.iters == 15
    assert np.array_equal(pose, far) and np.array_equal(po, far)


def test_empty_and_tiny_maps(engine):
    f, truth, guess = reg_case(7)
    tiny = np.zeros((3, 4), np.float32)
    mid = engine.map_create(tiny, tiny, gate_hint=1.0)   # Q6: < 5 map points => no correspondences
    pose, res, _ = engine.scan2map(mid, f["corner"], f["surf"], guess, E.lm_params("A"))
    assert res.status == E.FEW_CORRESPONDENCES
    assert np.array_equal(pose, np.asarray(guess, np.float32))
    engine.map_destroy(mid)
    mid = engine.map_create(np.zeros((0, 4), np.float32), np.zeros((0, 4), np.float32))
    pose, res, _ = engine.scan2map(mid, f["corner"], f["surf"], guess, E.lm_params("A"))
    assert res.status == E.FEW_CORRESPONDENCES
    engine.map_destroy(mid)


def test_degenerate_corridor_quirk_q1(engine):
    """A featureless ground plane leaves x, y, yaw unconstrained: isDegenerate at iteration 0, then the
    zeroed local matP makes X = 0 and the loop reports convergence at iteration 1 (quirk Q1)."""
    rng = np.random.default_rng(5)
    g = np.arange(-30, 30, 0.4, dtype=np.float32)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    ms = np.zeros((gx.size, 4), np.float32); ms[:, 0] = gx.ravel(); ms[:, 1] = gy.ravel(); ms[:, 2] = -1.73
    ms[:, :3] += rng.normal(0, 0.005, (len(ms), 3)).astype(np.float32)
    mc = ms[:10].copy()
    surf = np.zeros((3000, 4), np.float32)
    surf[:, :2] = rng.uniform(-20, 20, (3000, 2)); surf[:, 2] = -1.73 + rng.normal(0, 0.01, 3000)
    corner = surf[:20].copy()
    guess = np.array([0.003, -0.002, 0.01, 0.1, -0.1, 0.05], np.float32)
    mid = engine.map_create(mc, ms, gate_hint=1.0)
    pose_g, res_g, log_g = engine.scan2map(mid, corner, surf, guess, E.lm_params("A"), log=True)
    pose_o, res_o, log_o = orc.scan2map(corner, surf, mc, ms, guess, orc.lm_params("A"))
    engine.map_destroy(mid)
    assert res_o.is_degenerate == 1 and res_g.is_degenerate == 1
    assert res_o.iters == res_g.iters == 2 and res_g.converged == 1
    assert np.array(log_g[1].X).tolist() == [0.0] * 6
    er, et = synth.pose_error(pose_o, pose_g)
    assert er <= ROT_TOL and et <= TRANS_TOL


@pytest.mark.parametrize("variant", ["A", "B"])
def test_knn_check_path_is_bit_identical_to_full_search(monkeypatch, variant):
    """k_knn_check re-uses the previous iteration's neighbours when it can PROVE (safe radius vs movement) that they
    are still the exact 5-NN; LISREG_KNN_NOSKIP=1 searches every query from scratch.  Every per-iteration record
    (selection counts, A^T A, A^T b, step, pose) must be identical bit for bit, single and batched."""
    m = local_map()
    cases = [reg_case(s) for s in (0, 1, 2)]
    prm = E.lm_params(variant, early_exit=0, max_iters=8)
    logs = {}
    for noskip in ("1", "0"):
        monkeypatch.setenv("LISREG_KNN_NOSKIP", noskip)
        eng = E.Engine(device=0)
        mid = eng.map_create(m["corner"], m["surf"], gate_hint=2.0 if variant == "B" else 1.0)
        out = []
        for f, truth, guess in cases:
            pose, res, log = eng.scan2map(mid, f["corner"], f["surf"], guess, prm, clabel=f["corner_label"], slabel=f["surf_label"], log=True)
            out.append((pose.tobytes(), res.iters, res.n_sel_last,
                        [(bytes(bytearray(L.AtA)), bytes(bytearray(L.AtB)), bytes(bytearray(L.pose)), L.n_corner_sel, L.n_surf_sel) for L in log]))
        # the batched path tiles differently (512-query tiles): compare it with itself across the two modes
        items = [(mid, f["corner"], f["surf"], f["corner_label"], f["surf_label"]) for f, _, _ in cases] * 12
        poses, res, _ = eng.scan2map_batch(items, [g for _, _, g in cases] * 12, prm)
        out.append((np.asarray(poses).tobytes(), [r.n_sel_last for r in res]))
        logs[noskip] = out
        eng.close()
    assert logs["0"] == logs["1"]


def test_map_distance_filter_matches_oracle(engine, gpu_map):
    """Map-based dynamic removal (map_scan_feature_pts_distance_removal, subMap.h:1063-1098): keep mask bit-exact."""
    mid, m = gpu_map
    rng = np.random.default_rng(11)
    base = m["surf"][rng.choice(len(m["surf"]), 4000, replace=False)].copy()
    # displacements spanning every branch: on the map (< near), inside the dynamic band, beyond dyn_max, outside the disc
    disp = rng.choice([0.0, 0.01, 0.1, 0.5, 1.5, 4.0, 8.0], len(base))[:, None] * rng.standard_normal((len(base), 3))
    feat = base.copy(); feat[:, :3] += disp.astype(np.float32)
    for args in (dict(), dict(center_radius=15.0, dyn_min=0.5, dyn_max=2.0, near=0.05), dict(dyn_max=float(np.finfo(np.float32).max))):
        ko = orc.map_distance_filter(feat, m["surf"], **args)
        kg = engine.map_distance_filter(mid, 1, feat, **args)
        assert np.array_equal(ko, kg), args
        assert 0 < kg.sum() < len(kg)
    small = feat[:10]
    assert engine.map_distance_filter(mid, 1, small).all() and orc.map_distance_filter(small, m["surf"]).all()


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("gate", [1.0, 2.0])
def test_knn5_tie_order_on_lattice_map(engine, which, gate):
    """Exactly equidistant and duplicated map points (lattice + coinciding centroids, shuffled): the neighbour INDICES
    equal the oracle's - bit-equal distances are ordered by original index on both sides - through the block scan, the
    wide scan and the ball walk (k_knn5 alternates the deferred paths)."""
    m = lattice_map()
    mid = engine.map_create(m["corner"], m["surf"], gate_hint=gate)
    cloud = m["corner"] if which == 0 else m["surf"]
    q = lattice_queries(m, which)
    idx_o, sqd_o = orc.knn(cloud, q, 5)
    idx_g, sqd_g = engine.knn5(mid, which, q, gate)
    engine.map_destroy(mid)
    inside = sqd_o < gate
    assert (sqd_o[:, 3] == sqd_o[:, 4]).mean() > 0.1          # ties at the 5th place are the rule here
    assert np.array_equal(np.where(inside, idx_o, -1), idx_g)
    assert np.array_equal(sqd_o[inside], sqd_g[inside])


@pytest.mark.parametrize("variant", ["A", "B"])
def test_registration_on_lattice_map_matches_oracle(engine, variant):
    """Whole Gauss-Newton loop against a map full of ties: every search path (check / proof refresh, block scan, wide
    scan, warp-cooperative merge) must pick the oracle's neighbours, otherwise the fits - and the selection counts -
    drift.  Selection counts equal per iteration, pose within the north_star tolerance."""
    m = lattice_map()
    rng = np.random.default_rng(31)
    truth = np.array([0.004, -0.003, 0.3, 1.0, -0.5, 0.02], np.float32)
    T = synth.pose_to_T(truth)
    def scan_of(cloud, n):
        sel = np.sort(rng.choice(len(cloud), n, replace=False))
        w = cloud[sel, :3].astype(np.float64) + rng.normal(0, 0.004, (n, 3))
        out = np.zeros((n, 4), np.float32); out[:, :3] = ((w - T[:3, 3]) @ T[:3, :3]).astype(np.float32)
        return out
    corner = scan_of(m["corner"], 400); surf = scan_of(m["surf"], 4000)
    guess = synth.perturb_pose(truth, rng, rot=0.01, trans=0.1)
    gate = 1.0 if variant == "A" else 2.0
    mid = engine.map_create(m["corner"], m["surf"], gate_hint=gate)
    kw = dict(early_exit=0, max_iters=8)
    pose_o, res_o, log_o = orc.scan2map(corner, surf, m["corner"], m["surf"], guess, orc.lm_params(variant, **kw))
    pose_g, res_g, log_g = engine.scan2map(mid, corner, surf, guess, E.lm_params(variant, **kw), log=True)
    # batched path (512-query tiles, warp-cooperative late iterations)
    poses_b, res_b, _ = engine.scan2map_batch([(mid, corner, surf, None, None)] * 40, [guess] * 40, E.lm_params(variant, **kw))
    engine.map_destroy(mid)
    for lo, lg in zip(log_o, log_g):
        assert (lo.n_corner_sel, lo.n_surf_sel) == (lg.n_corner_sel, lg.n_surf_sel), (lo.n_corner_sel, lo.n_surf_sel, lg.n_corner_sel, lg.n_surf_sel)
        A_o, A_g = np.array(lo.AtA), np.array(lg.AtA)
        assert np.abs(A_o - A_g).max() <= 2e-4 * np.abs(A_o).max()
    assert log_o[0].n_sel > 2000
    er, et = synth.pose_error(pose_o, pose_g)
    assert er <= ROT_TOL and et <= TRANS_TOL, (er, et)
    for pb, rb in zip(poses_b, res_b):
        eb = synth.pose_error(pose_o, pb)
        assert eb[0] <= ROT_TOL and eb[1] <= TRANS_TOL and rb.n_sel_last == res_o.n_sel_last


def test_lattice_ties_at_iteration_zero(engine):
    """Guess = identity on scan points that sit on binary-exact offsets of the lattice: at iteration 0 most queries have
    several bit-equal neighbour distances (also at the 5th / 6th place), so the line / plane fits depend on the
    tie-break.  Selection counts equal, A^T A within the usual 2e-4 at every iteration, single and batched."""
    m = lattice_map()
    corner = lattice_queries(m, 0, n=400, seed=5); surf = lattice_queries(m, 1, n=4000, seed=6)
    corner[:, 0] += np.float32(1 / 128); surf[:, 2] += np.float32(1 / 128); surf[:, 1] += np.float32(1 / 256)   # off the lines / planes: no 0 / 0
    guess = np.zeros(6, np.float32)
    mid = engine.map_create(m["corner"], m["surf"], gate_hint=1.0)
    kw = dict(early_exit=0, max_iters=4)
    pose_o, res_o, log_o = orc.scan2map(corner, surf, m["corner"], m["surf"], guess, orc.lm_params("A", **kw))
    pose_g, res_g, log_g = engine.scan2map(mid, corner, surf, guess, E.lm_params("A", **kw), log=True)
    poses_b, res_b, _ = engine.scan2map_batch([(mid, corner, surf, None, None)] * 40, [guess] * 40, E.lm_params("A", **kw))
    engine.map_destroy(mid)
    for lo, lg in zip(log_o, log_g):
        assert (lo.n_corner_sel, lo.n_surf_sel) == (lg.n_corner_sel, lg.n_surf_sel)
        A_o, A_g = np.array(lo.AtA), np.array(lg.AtA)
        assert np.abs(A_o - A_g).max() <= 2e-4 * np.abs(A_o).max()
    er, et = synth.pose_error(pose_o, pose_g)
    assert er <= 1e-6 and et <= 1e-5, (er, et)
    for pb in poses_b:
        eb = synth.pose_error(pose_o, pb)
        assert eb[0] <= 1e-6 and eb[1] <= 1e-5


def test_non_finite_map_is_rejected(engine):
    """A NaN / Inf coordinate in a map cloud must fail loudly (LISREG_ERR_ARG) instead of planning a garbage grid; the
    context stays usable afterwards."""
    m = local_map()
    bad = m["surf"][:5000].copy(); bad[17, 0] = np.inf
    with pytest.raises(E.LisregError):
        engine.map_create(m["corner"][:1000], bad, gate_hint=1.0)
    bad[17, 0] = -np.inf
    with pytest.raises(E.LisregError):
        engine.map_create(bad, m["surf"][:1000], gate_hint=1.0)
    mid = engine.map_create(m["corner"][:1000], m["surf"][:5000], gate_hint=1.0)
    engine.map_destroy(mid)
