"""CPU known-answer tests of the de-skew restatement (oracle/orc_deskew.cpp; laserProcessing.cpp:368-462)."""
import numpy as np

from oracle import orc


def test_pure_yaw_rate_is_a_rotation_about_z():
    rng = np.random.default_rng(0)
    n = 500
    pts = np.zeros((n, 4), np.float32); pts[:, :3] = rng.uniform(-30, 30, (n, 3)); pts[:, 3] = rng.uniform(0, 100, n)
    time = np.sort(rng.uniform(0.0, 0.1, n)).astype(np.float32)
    t0 = 50.0
    imu_time = t0 - 0.01 + np.arange(40) * 0.004
    w = 0.8                                                        # rad/s about z
    imu_rot = np.zeros((40, 3)); imu_rot[:, 2] = w * (imu_time - imu_time[0])
    src = rng.permutation(n).astype(np.int32)                      # extraction order is arbitrary; the FIRST input index is the reference
    out = orc.deskew(pts, time, src, imu_time, imu_rot, t0)
    first = src.min()
    for k in (0, 17, n - 1):
        i = src[k]
        a = w * (float(time[i]) - float(time[first]))             # yaw accumulated since the reference point
        c, s = np.cos(a), np.sin(a)
        exp = np.array([c * pts[i, 0] - s * pts[i, 1], s * pts[i, 0] + c * pts[i, 1], pts[i, 2]])
        assert np.abs(out[k, :3] - exp).max() < 2e-4, (k, out[k], exp)
        assert out[k, 3] == pts[i, 3]
    k0 = int(np.where(src == first)[0][0])
    assert np.abs(out[k0, :3] - pts[first, :3]).max() < 1e-5       # the reference point does not move


def test_lookup_edges_and_disabled_table():
    pts = np.array([[10, 0, 0, 1], [0, 10, 0, 2], [0, 0, 10, 3]], np.float32)
    src = np.arange(3, dtype=np.int32)
    # disabled: pass-through
    assert np.array_equal(orc.deskew(pts, np.zeros(3, np.float32), src, np.zeros(0), np.zeros((0, 3)), 0.0), pts)
    # times before the first / after the last entry clamp to the end entries (findRotation :379-386)
    imu_time = np.array([1.0, 2.0, 3.0]); imu_rot = np.array([[0, 0, 0.0], [0, 0, 0.1], [0, 0, 0.3]])
    time = np.array([0.0, 1.5, 9.0], np.float32)                  # t = 0.5 (before), 2.0 (exactly an entry), 9.5 (after)
    out = orc.deskew(pts, time, src, imu_time, imu_rot, 0.5)
    # point 0: rot = entry 0 = 0 (reference);  point 1 at t = 2.0: pointTime < imuTime[2] -> front = 2, interpolation
    # between entries 1 and 2 with ratioFront = 0 -> 0.1 rad;  point 2: clamped to the last entry, 0.3 rad
    a1, a2 = 0.1, 0.3
    assert np.abs(out[0, :3] - pts[0, :3]).max() < 1e-6
    assert np.abs(out[1, :3] - np.array([-np.sin(a1) * 10, np.cos(a1) * 10, 0])).max() < 1e-5
    assert np.abs(out[2, :3] - pts[2, :3]).max() < 1e-6            # rotation about z leaves a point on the z axis alone
