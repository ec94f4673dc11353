"""BASELINE configs[4]-shaped streaming test: a VLP-16 (16 x 1800) sweep stream processed frame by frame through
the odometry flow (map = sliding window of keyframes + voxel grid, constant-velocity guess, keyframe rule), once
with the GPU engine and once with the CPU oracle behind the same host harness.  Per-frame pose tolerance:
1e-4 rad / 1e-3 m (north_star)."""
import numpy as np
import pytest

from lis_slam_b200 import engine as E
from lis_slam_b200 import stream, synth
from oracle import orc

from common import scene

pytestmark = pytest.mark.gpu


class OracleBackend:
    def __init__(self):
        self.maps = {}

    def extract_features(self, pts, ring, prm=None):
        return orc.extract_features(pts, ring, prm)

    def voxel_grid(self, pts, leaf):
        return orc.voxel_grid(pts, leaf)

    def map_create(self, corner, surf):
        self.maps[len(self.maps)] = (corner, surf)
        return len(self.maps) - 1

    def map_destroy(self, mid):
        self.maps[mid] = None

    def scan2map(self, mid, corner, surf, pose, prm):
        mc, ms = self.maps[mid]
        p, r, _ = orc.scan2map(corner, surf, mc, ms, pose, prm, log=False)
        return p, r


def test_vlp16_stream_matches_oracle(engine):
    sc = scene()
    n_frames = 8
    truth = [np.array([0.0, 0.0, 0.02 * np.sin(0.5 * t), -20.0 + 0.8 * t, 0.3 * np.sin(0.2 * t), 0.0], np.float32) for t in range(n_frames)]
    sweeps = [sc.scan(p, sensor="vlp16", seed=7000 + t) for t, p in enumerate(truth)]
    sg = stream.OdometryStream(stream.EngineBackend(engine), E.lm_params("A", surf_min_valid=100), E.feat_params(n_scan=16))
    so = stream.OdometryStream(OracleBackend(), orc.lm_params("A"), orc.feat_params(n_scan=16))
    for t, sw in enumerate(sweeps):
        pg = sg.push(sw["pts"], sw["ring"], initial_pose=truth[0])
        po = so.push(sw["pts"], sw["ring"], initial_pose=truth[0])
        er, et = synth.pose_error(po, pg)
        assert er <= 1e-4 and et <= 1e-3, (t, er, et)
    assert sg.keyframe_id == so.keyframe_id >= 3
    # the odometry tracks the ground truth (8 m/s forward motion) to centimetres (VLP-16 is sparse: 16 rings)
    er, et = synth.pose_error(truth[-1], sg.trajectory[-1])
    assert et < 0.3 and er < 0.03, (er, et)
    assert all((a is None) == (b is None) and (a is None or a.iters == b.iters) for a, b in zip(sg.results, so.results))


def _stream_traj(t, x0=-30.0):
    """8 m/s along x with a yaw wobble."""
    return np.array([0.0, 0.0, 0.02 * np.sin(0.05 * t), x0 + 0.8 * t, 0.3 * np.sin(0.02 * t), 0.0], np.float32)


def _run_device_vs_oracle_flow(eng, sensor, n_scan, n_frames, use_graph):
    """lisreg_odom_push (device-resident window, C-ABI) against stream.OdometryStream driving the CPU oracle."""
    sc = scene()
    oid = eng.odom_create(E.odom_params("A", n_scan=n_scan, use_graph=use_graph))
    so = stream.OdometryStream(OracleBackend(), orc.lm_params("A"), orc.feat_params(n_scan=n_scan))
    worst = (0.0, 0.0)
    n_kf_saved = 0
    for t in range(n_frames):
        sw = sc.scan(_stream_traj(t), sensor=sensor, seed=7000 + t, fast=True)
        pg, rg = eng.odom_push(oid, sw["pts"], sw["ring"], init_pose=_stream_traj(0))
        po = so.push(sw["pts"], sw["ring"], initial_pose=_stream_traj(0))
        er, et = synth.pose_error(po, pg)
        assert er <= 1e-4 and et <= 1e-3, (t, er, et)
        worst = (max(worst[0], er), max(worst[1], et))
        ro = so.results[-1]
        assert rg.keyframe_id == so.keyframe_id, (t, rg.keyframe_id, so.keyframe_id)
        if ro is not None:
            assert rg.lm.iters == ro.iters and rg.lm.status == ro.status, (t, rg.lm.iters, ro.iters)
        n_kf_saved += rg.keyframe_saved
    er, et = synth.pose_error(_stream_traj(n_frames - 1), pg)
    eng.odom_destroy(oid)
    assert n_kf_saved == so.keyframe_id >= 4
    return worst, (er, et)


@pytest.mark.parametrize("use_graph", [0, 1])
def test_odom_device_window_vlp16_matches_oracle_flow(use_graph):
    eng = E.Engine(device=0, own_stream=True)      # a non-default stream: the per-frame CUDA graph needs one
    worst, drift = _run_device_vs_oracle_flow(eng, "vlp16", 16, 14, use_graph)
    eng.close()
    assert drift[1] < 0.3 and drift[0] < 0.03


def test_odom_device_window_hdl64_64_frames_matches_oracle_flow():
    """BASELINE configs[1] shape: 64 HDL-64 frames (64 x 1800) through the device-resident sliding-window flow; every
    frame's pose within 1e-4 rad / 1e-3 m of the CPU oracle flow on identical input, same key-frame decisions and
    iteration counts.  The window fills up (>= 20 key frames would need more frames; 64 frames save ~35)."""
    eng = E.Engine(device=0, own_stream=True)
    worst, drift = _run_device_vs_oracle_flow(eng, "hdl64", 64, 64, 1)
    eng.close()
    assert drift[1] < 0.1 and drift[0] < 0.01


def test_odom_graph_replay_is_bit_identical_to_eager_launches():
    sc = scene()
    out = []
    for use_graph in (0, 1):
        eng = E.Engine(device=0, own_stream=True)
        oid = eng.odom_create(E.odom_params("A", n_scan=16, use_graph=use_graph))
        poses = []
        for t in range(10):
            sw = sc.scan(_stream_traj(t), sensor="vlp16", seed=7000 + t, fast=True)
            p, r = eng.odom_push(oid, sw["pts"], sw["ring"], init_pose=_stream_traj(0))
            poses.append(p.tobytes() + bytes([r.lm.iters, r.keyframe_saved]))
        out.append(poses)
        eng.odom_destroy(oid); eng.close()
    assert out[0] == out[1]


def _hint(t, rng, odom_on, imu_on):
    """cloud_info scalars an IMU pre-integration front end would publish: attitude = truth + noise, initial guess = the truth
    expressed from a drifting odom origin (only its increments matter, :331)."""
    tr = _stream_traj(t)
    imu_rpy = (tr[0] + 0.01 * rng.standard_normal(), tr[1] + 0.01 * rng.standard_normal(), tr[2] + 0.002 * rng.standard_normal())
    guess = (tr[3] + 30.0 + 0.02 * rng.standard_normal(), tr[4] + 0.02 * rng.standard_normal(), tr[5] + 0.01 * rng.standard_normal(),
             imu_rpy[0], imu_rpy[1], imu_rpy[2])
    return dict(imu_available=imu_on, odom_available=odom_on, imu_rpy=imu_rpy, initial_guess=guess)


@pytest.mark.parametrize("schedule", ["odom+imu", "imu_only", "odom_drops"])
def test_odom_cloud_info_hints_match_oracle_flow(schedule):
    """lisreg_odom_push_info: every branch of updateInitialGuess (:297-419 - first-frame IMU attitude, pre-integration
    increment, the fall-through IMU rotation increment on the frame that first sees odomAvailable, constant velocity when
    the odometry drops out) and the IMU slerp + clamps of transformUpdate (:976-1006), against the same flow in
    stream.OdometryStream around the CPU oracle.  Clamps are live (rot 0.002 rad, inside the IMU noise) so their order after the slerp matters."""
    eng = E.Engine(device=0, own_stream=True)
    sc = scene()
    prm = E.odom_params("A", n_scan=16, use_graph=1)
    prm.frame.lm.rot_tolerance, prm.frame.lm.z_tolerance = 0.002, 1000.0
    prm.imu_rpy_weight = 0.1                                                    # the shipped yaml value (params.yaml:88)
    oid = eng.odom_create(prm)
    tu = lambda pose, av, r, p, w: orc.transform_update(pose, av, np.float32(r), np.float32(p), w, 0.002, 1000.0)
    so = stream.OdometryStream(OracleBackend(), orc.lm_params("A", rot_tolerance=0.0, z_tolerance=0.0), orc.feat_params(n_scan=16),
                               imu_rpy_weight=0.1, transform_update=tu)
    rng = np.random.default_rng(99)
    used = set()
    for t in range(12):
        odom_on = {"odom+imu": t >= 1, "imu_only": False, "odom_drops": t in (2, 3, 4, 8, 9)}[schedule]
        h = _hint(t, rng, odom_on, True)
        sw = sc.scan(_stream_traj(t), sensor="vlp16", seed=7000 + t, fast=True)
        ci = E.cloud_info(h["imu_available"], h["odom_available"], h["imu_rpy"], h["initial_guess"])
        pg, rg = eng.odom_push_info(oid, sw["pts"], sw["ring"], ci)
        po = so.push(sw["pts"], sw["ring"], info=h)
        # the prediction itself (transformTobeMapped after updateInitialGuess) is host fp32 arithmetic: equal to rounding
        er, et = synth.pose_error(po, pg)
        assert er <= 1e-4 and et <= 1e-3, (schedule, t, er, et)
        ro = so.results[-1]
        assert rg.keyframe_id == so.keyframe_id
        if ro is not None:
            assert rg.lm.iters == ro.iters and rg.lm.status == ro.status, (t, rg.lm.iters, ro.iters)
            assert np.array_equal(np.array(rg.lm.pose[:], np.float32), pg)       # the result carries the updated pose
        if ro is not None and ro.status != 1:
            assert abs(pg[0]) <= np.float32(0.002) and abs(pg[1]) <= np.float32(0.002)
    expect = _stream_traj(11).copy(); expect[3] -= _stream_traj(0)[3]          # the hinted flow starts at the origin (:307-312)
    er, et = synth.pose_error(expect, pg)
    eng.odom_destroy(oid); eng.close()
    assert et < 0.3 and er < 0.05, (er, et)
