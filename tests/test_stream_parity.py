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
