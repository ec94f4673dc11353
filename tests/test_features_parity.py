"""GPU parity: LOAM feature extraction (F1-F5) through the C-ABI vs the CPU oracle, bit-exact
(integer/index work) on seeded synthetic HDL-64 / VLP-16 sweeps, plus the edge cases the reference
guards (empty sweep, duplicate cells = first hit wins, out-of-range rings, sparse rings)."""
import functools

import numpy as np
import pytest

from lis_slam_b200 import engine as E
from oracle import orc

from common import scene

pytestmark = pytest.mark.gpu

KEYS = ("src_index", "col_ind", "range", "start_ring", "end_ring", "curvature", "label",
        "corner_idx", "sharp_idx", "flat_idx", "surf_idx")


@functools.lru_cache(maxsize=None)
def sweep(seed, sensor="hdl64"):
    rng = np.random.default_rng(seed)
    pose = np.array([rng.uniform(-0.02, 0.02), rng.uniform(-0.02, 0.02), rng.uniform(-3, 3),
                     rng.uniform(-40, 40), rng.uniform(-2, 2), 0.0], np.float32)
    return scene().scan(pose, sensor=sensor, seed=2000 + seed)


def _compare(engine, pts, ring, po, pg):
    fo = orc.extract_features(pts, ring, po)
    fg = engine.extract_features(pts, ring, pg)
    assert fo["M"] == fg["M"]
    for k in KEYS:
        assert np.array_equal(fo[k], fg[k]), k
    return fo


@pytest.mark.parametrize("seed", [0, 1])
def test_hdl64_features_bit_exact(engine, seed):
    s = sweep(seed)
    fo = _compare(engine, s["pts"], s["ring"], orc.feat_params(), E.feat_params())
    assert len(fo["corner_idx"]) > 300 and len(fo["surf_idx"]) > 50000


def test_vlp16_and_downsample(engine):
    s = sweep(2, "vlp16")
    _compare(engine, s["pts"], s["ring"], orc.feat_params(n_scan=16), E.feat_params(n_scan=16))
    s = sweep(0)
    _compare(engine, s["pts"], s["ring"], orc.feat_params(downsample_rate=2), E.feat_params(downsample_rate=2))


def test_shuffled_input_first_hit_wins(engine):
    """Duplicates falling into one range-image cell: the first point in input order wins (:499)."""
    s = sweep(1)
    rng = np.random.default_rng(3)
    pts = np.concatenate([s["pts"], s["pts"][::3] * np.float32(1.002)])
    ring = np.concatenate([s["ring"], s["ring"][::3]])
    perm = rng.permutation(len(pts))
    fo = _compare(engine, pts[perm], ring[perm], orc.feat_params(), E.feat_params())
    assert fo["M"] <= 64 * 1800


def test_edge_cases(engine):
    po, pg = orc.feat_params(), E.feat_params()
    empty = np.zeros((0, 4), np.float32)
    fo = _compare(engine, empty, np.zeros(0, np.uint16), po, pg)
    assert fo["M"] == 0 and len(fo["surf_idx"]) == 0
    s = sweep(0)
    # rings out of range, ranges out of [min, max], a handful of points only
    pts = s["pts"][:2000].copy(); ring = s["ring"][:2000].copy()
    ring[::7] = 200
    pts[::11, :3] *= 100.0
    _compare(engine, pts, ring, po, pg)
    # only two rings populated: the other rings have start > end (skipped segments)
    keep = (s["ring"] == 10) | (s["ring"] == 40)
    _compare(engine, s["pts"][keep], s["ring"][keep], po, pg)


def test_deskew_matches_oracle(engine):
    """Motion de-skew of the projection (deskewPoint, laserProcessing.cpp:427-462): the de-skewed extracted cloud is
    bit-exact vs the oracle; range / column / every feature list are those of the ORIGINAL sweep (as upstream)."""
    from common import scene
    pose = np.array([0.01, -0.02, 0.7, 3.0, -1.0, 0.0], np.float32)
    s = scene().scan(pose, seed=2100)
    rng = np.random.default_rng(21)
    t_scan = 1000.25
    n_imu = 30
    imu_time = t_scan - 0.008 + np.arange(n_imu) * 0.0045 + rng.uniform(0, 1e-4, n_imu)      # covers [t_scan, t_scan + 0.1]
    rate = np.array([0.05, -0.03, 0.6])                                                      # rad/s: a turn
    imu_rot = np.zeros((n_imu, 3))
    for i in range(1, n_imu):
        imu_rot[i] = imu_rot[i - 1] + (rate + 0.02 * rng.standard_normal(3)) * (imu_time[i] - imu_time[i - 1])
    base = engine.extract_features(s["pts"], s["ring"])
    g = engine.extract_features(s["pts"], s["ring"], time=s["time"], imu_time=imu_time, imu_rot=imu_rot, time_scan_cur=t_scan)
    for k in ("src_index", "col_ind", "range", "curvature", "label", "corner_idx", "sharp_idx", "flat_idx", "surf_idx"):
        assert np.array_equal(base[k], g[k]), k
    o = orc.deskew(s["pts"], s["time"], g["src_index"], imu_time, imu_rot, t_scan)
    assert np.array_equal(o, g["ext_pts"])
    moved = np.linalg.norm(g["ext_pts"][:, :3] - s["pts"][g["src_index"], :3], axis=1)
    assert moved.max() > 0.5 and moved.min() < 1e-3          # 0.06 rad over the sweep at up to 70 m; the first point does not move
    # disabled table (deskewFlag -1 / IMU unavailable): points pass through
    g0 = engine.extract_features(s["pts"], s["ring"], time=s["time"], imu_time=np.zeros(0), imu_rot=np.zeros((0, 3)), time_scan_cur=t_scan)
    assert np.array_equal(g0["ext_pts"], s["pts"][g0["src_index"]])
    # a single table entry and point times outside the table: clamped to the end entries like findRotation
    g1 = engine.extract_features(s["pts"], s["ring"], time=s["time"], imu_time=imu_time[:3], imu_rot=imu_rot[:3], time_scan_cur=t_scan)
    o1 = orc.deskew(s["pts"], s["time"], g1["src_index"], imu_time[:3], imu_rot[:3], t_scan)
    assert np.array_equal(o1, g1["ext_pts"])


def _pack_layout(sw, which):
    """The same sweep in another memory layout (lisreg_cloud_layout presets)."""
    pts, ring, t = sw["pts"], sw["ring"], sw["time"]
    if which == E.LAYOUT_XYZ_RING or which == E.LAYOUT_XYZ_SYNTH_RING:
        return np.ascontiguousarray(pts[:, :3]), (ring if which == E.LAYOUT_XYZ_RING else None)
    rec = np.zeros(len(pts), dtype=np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                                             "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                                             "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32}))    # PCL PointXYZIRT (common.h:12-23)
    rec["x"], rec["y"], rec["z"], rec["intensity"], rec["ring"], rec["time"] = pts[:, 0], pts[:, 1], pts[:, 2], pts[:, 3], ring, t
    return rec, None


@pytest.mark.parametrize("sensor,n_scan", [("hdl64", 64), ("vlp16", 16)])
def test_cloud_layouts_give_identical_features(engine, sensor, n_scan):
    """PointCloud2-style layouts (32-byte PCL PointXYZIRT records read in place, bare xyz + ring array, xyz with the
    ring synthesised from the elevation angle as laserPretreatmentNode.cpp:95-126) produce exactly the index lists of the
    packed float4 + ring-array input (and therefore of the oracle).  Ring synthesis is checked on the VLP-16 shape, whose
    2 degree spacing is the one the upstream formula encodes (the synthetic HDL-64 uses a linear elevation table)."""
    sw = scene().scan(np.array([0.01, -0.01, 0.4, 3.0, 0.5, 0.0], np.float32), sensor=sensor, seed=123, fast=True)
    base = engine.extract_features(sw["pts"], sw["ring"], E.feat_params(n_scan=n_scan))
    ref = orc.extract_features(sw["pts"], sw["ring"], orc.feat_params(n_scan=n_scan))
    assert np.array_equal(base["surf_idx"], ref["surf_idx"]) and np.array_equal(base["corner_idx"], ref["corner_idx"])
    layouts = [E.LAYOUT_XYZ_RING, E.LAYOUT_PCL_XYZIRT] + ([E.LAYOUT_XYZ_SYNTH_RING] if sensor == "vlp16" else [])
    for which in layouts:
        prm = E.feat_params(n_scan=n_scan); prm.layout = E.cloud_layout(which)
        data, ring = _pack_layout(sw, which)
        out = engine.extract_features(data, ring, prm)
        for k in ("src_index", "col_ind", "range", "corner_idx", "sharp_idx", "flat_idx", "surf_idx", "curvature", "label", "start_ring", "end_ring"):
            assert np.array_equal(out[k], base[k]), (which, k)
    # bad layouts fail loudly
    prm = E.feat_params(n_scan=n_scan); prm.layout = E.cloud_layout(E.LAYOUT_XYZ_RING); prm.layout.off_y = 5
    with pytest.raises(E.LisregError):
        engine.extract_features(np.ascontiguousarray(sw["pts"][:, :3]), sw["ring"], prm)


def test_fused_front_end_equals_range_image_path():
    """k_feat_front (projection + compaction with the range-image slice in shared memory, look-back over ring groups) writes
    exactly what the global range-image kernels write: every output array of a single sweep (64 one-ring groups)
    is identical with and without LISREG_FEAT_FUSED=1 (batches: tests/test_frames_parity.py and the all-frames pose check of bench.py)."""
    import os
    from common import scene
    sc = scene()
    sweeps = [sc.scan(np.array([0.01 * k, -0.02, 0.1 * k, 1.0 * k, -0.5 * k, 0.0], np.float32), seed=3100 + k, fast=True) for k in range(3)]
    outs = {}
    for mode in ("0", "1"):
        os.environ["LISREG_FEAT_FUSED"] = mode
        try:
            eng = E.Engine(device=0)
            outs[mode] = [eng.extract_features(s["pts"], s["ring"]) for s in sweeps]
            eng.close()
        finally:
            os.environ.pop("LISREG_FEAT_FUSED", None)
    for a, b in zip(outs["0"], outs["1"]):
        assert set(a) == set(b)
        for k in a:
            assert np.array_equal(a[k], b[k]), k
