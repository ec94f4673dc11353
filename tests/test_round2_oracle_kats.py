"""CPU known-answer tests that pin the ROUND-2 parts of the oracle (pre-treatment, constant-velocity de-skew, local-map
insert / extract) against independent restatements written from the reference source on scipy / numpy - the GPU parity
tests compare the engine with this oracle, these tests keep the oracle itself honest."""
import numpy as np
from scipy.spatial.transform import Rotation

from oracle import orc


def test_constant_velocity_deskew_matches_scipy_rotations():
    """DistortionAdjust::AdjustCloud / UpdateMatrix (distortionAdjust.cpp:419-479): point k >= 1 becomes
    Rz(wz t) Ry(wy t) Rx(wx t) p + v t with t = time - period / 2; the first point is dropped; intensity is kept."""
    rng = np.random.default_rng(3)
    n = 2000
    p = np.zeros((n, 4), np.float32); p[:, :3] = rng.uniform(-40, 40, (n, 3)); p[:, 3] = rng.uniform(0, 255, n)
    t = np.sort(rng.uniform(0, 0.1, n)).astype(np.float32)
    v = np.array([8.0, -0.5, 0.1], np.float32); w = np.array([0.02, -0.03, 0.6], np.float32)
    out = orc.deskew_cv(p, t, 0.1, v, w)
    assert out.shape == (n - 1, 4) and np.array_equal(out[:, 3], p[1:, 3])
    rt = (t[1:] - np.float32(0.1) / np.float32(2.0)).astype(np.float64)
    ang = w.astype(np.float64)[None, :] * rt[:, None]
    R = Rotation.from_euler("ZYX", np.stack([ang[:, 2], ang[:, 1], ang[:, 0]], 1))     # intrinsic Z-Y-X = Rz * Ry * Rx
    exp = R.apply(p[1:, :3].astype(np.float64)) + v.astype(np.float64)[None, :] * rt[:, None]
    assert np.abs(out[:, :3] - exp).max() < 2e-5                                      # fp32 Eigen arithmetic vs fp64


def test_pretreatment_rings_and_times_of_a_constructed_sweep():
    """laserPretreatmentNode.cpp:60-230 on a sweep built to order: points exactly on the 16 VLP elevation angles (-15 .. +15
    in 2-degree steps) get ring (angle + 15) / 2, points closer than min_range are dropped, and a clockwise sweep of
    ascending azimuth gets a relative time that grows with the swept fraction of the revolution (0 .. period)."""
    H, n_scan, period = 360, 16, 0.1
    elev = np.deg2rad(-15.0 + 2.0 * np.arange(n_scan))
    az = -np.linspace(0.0, 2 * np.pi, H, endpoint=False) + 0.3          # velodyne spins clockwise: azimuth decreases
    pts = []
    for a in az:
        for r_id, e in enumerate(elev):
            rr = 10.0 + 0.5 * r_id
            pts.append([rr * np.cos(e) * np.cos(a), rr * np.cos(e) * np.sin(a), rr * np.sin(e), float(r_id)])
    pts = np.array(pts, np.float32)
    pts[5, :3] = [0.1, 0.1, 0.0]                                       # inside min_range: removed
    pts[9, :3] = np.nan                                                # removeNaNFromPointCloud
    out, ring, t = orc.pretreat(pts, n_scan, scan_period=period, min_range=1.0, max_range=70.0)
    assert len(out) == len(pts) - 2
    assert np.array_equal(ring, out[:, 3].astype(np.uint16))           # intensity carried the true ring id
    assert t.min() >= -1e-6 and t.max() <= period * 1.0001         # (fp32 rounding of the first column: -5e-10)
    # time is monotone in the firing order up to the ring interleave (all 16 rings of a column share the azimuth)
    col_t = np.array([t[ring == 3][k] for k in range(0, (ring == 3).sum(), 20)])
    assert np.all(np.diff(col_t) > 0)
    # the orientation span is measured from the first to the LAST point (endOri), so the last column closes the period, and
    # a column half way round sits at half the period
    assert abs(t[-1] - period) < 1e-3 * period and abs(t[ring == 3][(ring == 3).sum() // 2] - period / 2) < 0.01 * period


def test_local_map_insert_moves_clouds_and_extract_crops_to_the_box():
    """SubMapManager::insert_local_map + extractSlidingCloud (subMap.h:785-1055, subMapOptmizationNode.cpp:1369-1432):
    inserting with a pose moves every class cloud by pcl::getTransformation(pose) (checked against scipy), the bound is the
    exact extrema of all five clouds, and the extraction keeps only voxel centroids strictly inside the sensor box
    (+-70, +-70, -10 .. 20 moved with the current pose) intersected with the bound and padded by 2 m."""
    rng = np.random.default_rng(8)
    clouds = []
    for c in range(5):
        q = np.zeros((3000, 4), np.float32); q[:, :3] = rng.uniform(-90, 90, (3000, 3)) * np.array([1, 1, 0.15]); q[:, 3] = c
        clouds.append(q)
    pose = np.array([0.02, -0.01, 0.4, 3.0, -2.0, 0.5], np.float32)     # roll pitch yaw x y z
    sm = orc.Submap()
    counts = sm.insert(clouds, pose)
    assert counts == [3000] * 5
    R = Rotation.from_euler("ZYX", [pose[2], pose[1], pose[0]])
    for c in range(5):
        exp = R.apply(clouds[c][:, :3].astype(np.float64)) + pose[3:6].astype(np.float64)
        got = sm.get(c)
        assert np.abs(got[:, :3] - exp).max() < 5e-5 and np.array_equal(got[:, 3], clouds[c][:, 3])
    allp = np.concatenate([sm.get(c)[:, :3] for c in range(5)])
    assert np.array_equal(sm.bound, np.concatenate([allp.min(0), allp.max(0)]).astype(np.float64))
    cur = np.array([0.0, 0.0, 0.0, 3.0, -2.0, 0.5], np.float32)
    corner, surf, cnt = sm.extract(cur)
    # transform_bbx moves the box with its centre; get_intersection_bbx pads the intersection by 2 m (subMap.h:173-228)
    box_lo = np.maximum(np.array([3.0 - 70, -2.0 - 70, 0.5 - 10]), sm.bound[:3]) - 2.0
    box_hi = np.minimum(np.array([3.0 + 70, -2.0 + 70, 0.5 + 20]), sm.bound[3:]) + 2.0
    for cloud in (corner, surf):
        assert len(cloud) and np.all(cloud[:, :3] > box_lo) and np.all(cloud[:, :3] < box_hi)
    assert np.any(allp[:, 0] > box_hi[0]) and np.any(allp[:, 2] < box_lo[2])              # there was something to crop
    assert len(corner) == cnt[1] and len(surf) == cnt[0] + cnt[2] + cnt[3]                # pole -> corner map; ground + building + dynamic -> surface map
    assert sum(cnt) < 15000                                                                # the crop and the voxel filter removed points
    sm.close()


def test_loop_verify_recovers_a_known_key_frame_pose():
    """detectLoopClosureForSubMap (subMapOptmizationNode.cpp:2739-2916): the key-frame cloud is the submap's own geometry seen
    from a KNOWN pose T_true; the verifier starts from a pose that is 25 cm / 0.6 deg off (pose-based initial alignment
    submap^-1 * key, :2807-2809), ICP pulls it back, and with an identity relative pose the returned tCorrect - hence the
    loop constraint (:2876-2896) - is T_true: x, y, z and roll, pitch, yaw come back within a millimetre / 1e-4 rad."""
    from lis_slam_b200 import synth
    sc = synth.Scene(seed=1001)
    m = sc.sample_map(n_edge=0, n_surf=60000, seed=3001)["surf"]
    classes = [m[0::4], m[1::4], m[2::4], m[3::4], np.zeros((0, 4), np.float32)]
    sm = orc.Submap()
    sm.insert(classes, np.zeros(6, np.float32))                                       # submap frame = map frame
    true6 = np.array([0.01, -0.02, 0.3, 4.0, -1.5, 0.2], np.float32)                  # roll pitch yaw x y z
    R = Rotation.from_euler("ZYX", [true6[2], true6[1], true6[0]])
    sub = m[::5]
    key = sub.copy()
    key[:, :3] = R.inv().apply(sub[:, :3].astype(np.float64) - true6[3:6].astype(np.float64)).astype(np.float32)   # T_true * key = sub
    wrong6 = true6 + np.array([0.0, 0.0, 0.01, 0.2, -0.15, 0.0], np.float32)
    cand = dict(submap=sm, use_epsc=False, prekey_pose6=np.zeros(6, np.float32), epsc_T=np.eye(4, dtype=np.float32), submap_pose6=np.zeros(6, np.float32))
    r = orc.loop_verify(key, wrong6, np.zeros(6, np.float32), [cand])
    assert r["found"] == 1 and r["best"] == 0 and r["converged"][0] == 1 and r["fitness"][0] < 1e-3
    c = r["constraint6"]                                                              # x y z roll pitch yaw of tCorrect
    assert np.abs(c[:3] - true6[3:6]).max() < 2e-3
    assert np.abs(c[3:] - true6[:3]).max() < 2e-4
    # the correction is what separates the wrong start from the truth: correction * key2pre == tCorrect
    assert np.abs(r["correction"] @ r["key2pre"] - r["t_correct"]).max() < 1e-5
    # a fitness threshold below the achieved score rejects the loop but still reports the best candidate
    r2 = orc.loop_verify(key, wrong6, np.zeros(6, np.float32), [cand], fitness_threshold=0.0)
    assert r2["found"] == 0 and r2["best"] == 0
    sm.close()
