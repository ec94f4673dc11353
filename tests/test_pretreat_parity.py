"""GPU parity of the sweep pre-treatment (SURVEY.md 8f "next" #3) against the CPU oracle (oracle/orc_pretreat.cpp):
ring / time synthesis (laserPretreatmentNode.cpp:60-230) and the constant-velocity de-skew
(DistortionAdjust::AdjustCloud, distortionAdjust.cpp:419-479) are per-point float / index work -> bit-exact; the de-skew
inside the batched frame pipeline (lisreg_frame_params.deskew on PointXYZIRT records) == the single-sweep de-skew path
chained with the voxel grid and the registration."""
import ctypes as C

import numpy as np
import pytest

from lis_slam_b200 import engine as E
from lis_slam_b200 import synth
from oracle import orc

from common import local_map, scene

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sensor,n_scan", [("vlp16", 16), ("hdl64", 64)])
def test_ring_time_synthesis_bit_exact(engine, sensor, n_scan):
    sw = scene().scan(np.array([0.01, 0.0, 0.7, 3.0, 0.5, 0.0], np.float32), sensor=sensor, seed=11, fast=True)
    pts = sw["pts"].copy()
    pts[5] = np.nan; pts[100, 1] = np.inf; pts[200, :3] = 0.01          # removeNaN / removeClosedPointCloud cases
    for rng_max in (70.0, 25.0):
        po, ro, to = orc.pretreat(pts, n_scan, 0.1, 1.0, rng_max)
        pg, rg, tg = engine.pretreat(pts, n_scan, 0.1, 1.0, rng_max)
        assert len(po) == len(pg) > 1000
        assert np.array_equal(po, pg) and np.array_equal(ro, rg) and np.array_equal(to, tg)
    if sensor == "vlp16":        # 2 degree spacing = the upstream formula: the synthesised ids are the generator's
        keep = np.isfinite(pts[:, :3]).all(1) & ((pts[:, :3] ** 2).sum(1) >= 1.0) & ((pts[:, :3] ** 2).sum(1) <= 70.0 ** 2)
        po, ro, to = orc.pretreat(pts, n_scan, 0.1, 1.0, 70.0)
        assert np.array_equal(ro, sw["ring"][keep])
    # empty / everything filtered
    p0, r0, t0 = engine.pretreat(np.zeros((0, 4), np.float32), n_scan)
    assert len(p0) == 0
    p1, r1, t1 = engine.pretreat(pts[:50], n_scan, 0.1, 500.0, 600.0)
    assert len(p1) == 0 == len(orc.pretreat(pts[:50], n_scan, 0.1, 500.0, 600.0)[0])


def test_constant_velocity_deskew_bit_exact(engine):
    sw = scene().scan(np.array([0.0, 0.0, 0.2, -5.0, 0.3, 0.0], np.float32), sensor="hdl64", seed=12, fast=True)
    for lin, ang in (((8.0, 0.3, -0.1), (0.01, -0.02, 0.3)), ((0.0, 0.0, 0.0), (0.0, 0.0, 0.0)), ((-2.0, 1.0, 0.0), (0.0, 0.0, -1.2))):
        o = orc.deskew_cv(sw["pts"], sw["time"], 0.1, lin, ang)
        g = engine.deskew_cv(sw["pts"], sw["time"], 0.1, lin, ang)
        assert o.shape == g.shape == (len(sw["pts"]) - 1, 4)
        assert np.array_equal(o, g)
    assert engine.deskew_cv(sw["pts"][:1], sw["time"][:1], 0.1, (1, 0, 0), (0, 0, 1)).shape == (0, 4)


def _xyzirt(sw):
    rec = np.zeros(len(sw["pts"]), dtype=np.dtype({"names": ["x", "y", "z", "intensity", "ring", "time"],
                                                   "formats": ["<f4", "<f4", "<f4", "<f4", "<u2", "<f4"],
                                                   "offsets": [0, 4, 8, 16, 20, 24], "itemsize": 32}))
    p = sw["pts"]
    rec["x"], rec["y"], rec["z"], rec["intensity"], rec["ring"], rec["time"] = p[:, 0], p[:, 1], p[:, 2], p[:, 3], sw["ring"], sw["time"]
    return rec


def test_deskew_inside_the_batched_frame_pipeline(engine):
    """lisreg_frames_batch_arena on PointXYZIRT records with one IMU rotation table per frame == per frame:
    lisreg_extract_features_deskew -> voxel grid of the de-skewed corner / surface clouds -> lisreg_scan2map."""
    m = local_map()
    mid = engine.map_create(m["corner"], m["surf"], gate_hint=1.0)
    rng = np.random.default_rng(3)
    frames = []
    for k in range(3):
        truth = synth.random_pose(np.random.default_rng(600 + k)); guess = synth.perturb_pose(truth, rng)
        sw = scene().scan(truth, sensor="hdl64", seed=2100 + k, fast=True)
        t_scan = 1000.0 + 0.1 * k
        n_imu = 30 if k != 1 else 0                                   # frame 1: no IMU table -> points pass through
        imu_time = t_scan - 0.008 + np.arange(n_imu) * 0.0045 + rng.uniform(0, 1e-4, n_imu)
        imu_rot = np.zeros((n_imu, 3))
        for i in range(1, n_imu):
            imu_rot[i] = imu_rot[i - 1] + (np.array([0.02, -0.01, 0.15]) + 0.02 * rng.standard_normal(3)) * (imu_time[i] - imu_time[i - 1])
        frames.append((sw, guess, t_scan, np.ascontiguousarray(imu_time), np.ascontiguousarray(imu_rot)))
    # ---- batched: records in one arena, deskew entries per frame ----
    recs = [_xyzirt(sw) for sw, *_ in frames]
    arena = np.concatenate([r.view(np.uint8).reshape(-1) for r in recs])
    items = (E.FrameItem * 3)(); off = 0
    for i, r in enumerate(recs):
        items[i] = E.FrameItem(off, 0, len(r), mid); off += r.nbytes
    dsk = (E.Deskew * 3)()
    for i, (_, _, t_scan, it, ir) in enumerate(frames):
        dsk[i] = E.Deskew(it.ctypes.data if len(it) else None, ir.ctypes.data if len(it) else None, len(it), 0, t_scan)
    prm = E.frame_params("A"); prm.feat.layout = E.cloud_layout(E.LAYOUT_PCL_XYZIRT); prm.deskew = C.cast(dsk, C.c_void_p)
    pose = np.stack([g for _, g, *_ in frames]).astype(np.float32); res = (E.LmResult * 3)()
    engine.frames_batch_arena(items, 3, arena.ctypes.data, arena.nbytes, pose, prm, res)
    # ---- reference chain through the single-sweep entry points ----
    for i, (sw, guess, t_scan, it, ir) in enumerate(frames):
        g = engine.extract_features(sw["pts"], sw["ring"], time=sw["time"], imu_time=it, imu_rot=ir, time_scan_cur=t_scan)
        ext = g["ext_pts"]
        if len(it):
            assert np.abs(ext[:, :3] - sw["pts"][g["src_index"], :3]).max() > 1e-3        # the de-skew really moved points
        corner = engine.voxel_grid(np.ascontiguousarray(ext[g["corner_idx"]]), 0.2)
        surf = engine.voxel_grid(np.ascontiguousarray(ext[g["surf_idx"]]), 0.4)
        p1, r1, _ = engine.scan2map(mid, corner, surf, guess, E.lm_params("A"))
        assert (res[i].n_corner, res[i].n_surf) == (len(corner), len(surf))
        assert res[i].iters == r1.iters
        er, et = synth.pose_error(p1, pose[i])
        assert er <= 1e-6 and et <= 1e-5, (i, er, et)       # same clouds; only the tile order of the fp64 sums differs (single vs batch tiling)
    engine.map_destroy(mid)
