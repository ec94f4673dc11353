"""Streaming (per-frame) harness for BASELINE configs[1] / configs[4]: the host-side flow of
OdomEstimationNode::laserCloudInfoHandler (src/node/odomEstimationNode.cpp:163-239) around the device calls.

Per frame: constant-velocity initial guess (updateInitialGuess, :353-391) -> local map = concatenation of the last
<= 19 keyframe clouds (already in the map frame, newest first) -> voxel grid 0.2 / 0.4 m (:185-207) -> voxel grid of
the frame's corner / surface clouds (currentCloudInit, :260-281) -> scan2SubMapOptimization (:596-626) -> keyframe
rule (:216-229) and saveKeyFrames (:421-478).  Frame t needs the pose and map of frame t-1, so this mode does not
shard: multi-GPU = independent replicas.

`backend` provides the compute: extract_features(pts, ring), voxel_grid(pts, leaf), map_create(corner, surf),
map_destroy(id), scan2map(map_id, corner, surf, pose6, params).  lis_slam_b200.engine.Engine satisfies it (the
product); the tests also drive it with an adapter around the CPU oracle to compare trajectories.
"""
import numpy as np

from . import synth

KEYFRAME_MIN_DISTANCE = 1.4   # keyFrameMiniDistance, config/params.yaml
KEYFRAME_MIN_YAW = 0.5        # keyFrameMiniYaw
WINDOW = 19                   # while (laserCloudSurfVec.size() >= 20) erase(begin)  (:463-467)


f32 = np.float32


def _T16(pose6):
    """pcl::getTransformation in fp32 (same closed form as csrc/lm.cuh state_refresh / lisreg.cu odom_T16)."""
    p = np.asarray(pose6, f32)
    c = lambda a: f32(np.cos(np.float64(a)))
    s_ = lambda a: f32(np.sin(np.float64(a)))
    A, B, Cc, D, E, F = c(p[2]), s_(p[2]), c(p[1]), s_(p[1]), c(p[0]), s_(p[0])
    DE, DF = f32(D * E), f32(D * F)
    T = np.zeros((4, 4), f32)
    T[0] = [f32(A * Cc), f32(f32(A * DF) - f32(B * E)), f32(f32(B * F) + f32(A * DE)), p[3]]
    T[1] = [f32(B * Cc), f32(f32(A * E) + f32(B * DF)), f32(f32(B * DE) - f32(A * F)), p[4]]
    T[2] = [f32(-D), f32(Cc * F), f32(Cc * E), p[5]]
    T[3, 3] = 1
    return T


def _cof(T, i, j):
    i1, i2, j1, j2 = (i + 1) % 3, (i + 2) % 3, (j + 1) % 3, (j + 2) % 3
    return f32(f32(T[i1, j1] * T[i2, j2]) - f32(T[i1, j2] * T[i2, j1]))


def _inv(T):
    """Eigen::Affine3f::inverse(): cofactor inverse of the linear part, -linear^-1 * t (lisreg.cu odom_inv)."""
    c0, c1, c2 = _cof(T, 0, 0), _cof(T, 1, 0), _cof(T, 2, 0)
    det = f32(f32(f32(c0 * T[0, 0]) + f32(c1 * T[1, 0])) + f32(c2 * T[2, 0]))
    invdet = f32(f32(1) / det)
    R = np.zeros((4, 4), f32)
    for i in range(3):
        for j in range(3):
            R[i, j] = f32(_cof(T, j, i) * invdet)
    for i in range(3):
        R[i, 3] = f32(f32(f32(f32(-R[i, 0]) * T[0, 3]) + f32(f32(-R[i, 1]) * T[1, 3])) + f32(f32(-R[i, 2]) * T[2, 3]))
    R[3, 3] = 1
    return R


def _mul(A, B):
    Cm = np.zeros((4, 4), f32)
    for i in range(4):
        for j in range(4):
            a = f32(0)
            for k in range(4):
                a = f32(a + f32(A[i, k] * B[k, j]))
            Cm[i, j] = a
    return Cm


def _euler(T):
    """pcl::getTranslationAndEulerAngles -> [roll, pitch, yaw, x, y, z] fp32."""
    return np.array([f32(np.arctan2(np.float64(T[2, 1]), np.float64(T[2, 2]))), f32(np.arcsin(np.float64(-T[2, 0]))),
                     f32(np.arctan2(np.float64(T[1, 0]), np.float64(T[0, 0]))), T[0, 3], T[1, 3], T[2, 3]], f32)


def transform_cloud(pts4, pose6):
    """common.cpp:113-150 transformPointCloud(cloudIn, PointTypePose*): q = R p + t in fp32, intensity kept."""
    T = _T16(pose6)
    out = np.array(pts4, dtype=np.float32, copy=True)
    x, y, z = pts4[:, 0], pts4[:, 1], pts4[:, 2]
    out[:, 0] = T[0, 0] * x + T[0, 1] * y + T[0, 2] * z + T[0, 3]
    out[:, 1] = T[1, 0] * x + T[1, 1] * y + T[1, 2] * z + T[1, 3]
    out[:, 2] = T[2, 0] * x + T[2, 1] * y + T[2, 2] * z + T[2, 3]
    return out


class OdometryStream:
    def __init__(self, backend, lm_params, feat_params=None, corner_leaf=0.2, surf_leaf=0.4, use_imu_heading=False,
                 imu_rpy_weight=0.01, transform_update=None):
        self.be, self.prm, self.fprm = backend, lm_params, feat_params
        self.corner_leaf, self.surf_leaf = corner_leaf, surf_leaf
        self.pose = np.zeros(6, np.float32)          # transformTobeMapped
        self.last_pose = None                        # lastTransformTobeMapped
        self.key_pose = np.zeros(6, np.float32)      # transformPriFrame
        self.kf_corner, self.kf_surf = [], []
        self.keyframe_id = 0
        self.first = True
        self.first_trans = False
        self.deltaR, self.deltaT = f32(100), f32(100)     # members (:70-71): keep the last solved step across frames
        self.use_imu_heading = use_imu_heading            # useImuHeadingInitialization
        self.imu_rpy_weight = imu_rpy_weight              # imuRPYWeight
        self.transform_update = transform_update          # callable(pose6, imu_available, imu_roll, imu_pitch, weight, rot_tol, z_tol) -> pose6
        self.last_imu = np.eye(4, dtype=f32)              # lastImuTransformation
        self.last_imu_pre = None                          # lastImuPreTransformation (None = lastImuPreTransAvailable false)
        self.trajectory, self.results = [], []

    def _update_initial_guess(self, initial_pose, info=None):
        """updateInitialGuess (:297-419), Affine3f arithmetic in fp32.  info = None: no hints (odomAvailable = imuAvailable =
        false; first pose = initial_pose); else a dict with imu_available, odom_available, imu_rpy (3), initial_guess (x, y, z,
        roll, pitch, yaw) - the scalar fields of lis_slam::cloud_info."""
        if not self.first_trans:
            if info is not None:
                r, p, y = (f32(v) for v in info["imu_rpy"])
                self.pose = np.array([r, p, y if self.use_imu_heading else 0, 0, 0, 0], f32)
                self.last_imu = _T16([r, p, y, 0, 0, 0])
            else:
                self.pose = np.zeros(6, f32) if initial_pose is None else np.asarray(initial_pose, f32).copy()
            self.first_trans = True
            return
        odom_av = info is not None and bool(info["odom_available"])
        if odom_av:
            g = np.asarray(info["initial_guess"], f32)
            T_back = _T16([g[3], g[4], g[5], g[0], g[1], g[2]])
            if self.last_imu_pre is None:
                self.last_imu_pre = T_back                      # and fall through to the imuAvailable block (:327-330)
            else:
                incre = _mul(_inv(self.last_imu_pre), T_back)
                self.pose = _euler(_mul(_T16(self.pose), incre))
                self.last_imu_pre = T_back
                r, p, y = (f32(v) for v in info["imu_rpy"])
                self.last_imu = _T16([r, p, y, 0, 0, 0])
                return
        if not odom_av:
            if self.last_pose is None:
                self.last_pose = self.pose.copy()
                return
            T_back, T_last = _T16(self.pose), _T16(self.last_pose)
            self.last_pose = self.pose.copy()
            incre = _mul(_inv(T_last), T_back)
            self.pose = _euler(_mul(T_back, incre))
            return
        if info["imu_available"]:
            r, p, y = (f32(v) for v in info["imu_rpy"])
            T_back = _T16([r, p, y, 0, 0, 0])
            incre = _mul(_inv(self.last_imu), T_back)
            self.pose = _euler(_mul(_T16(self.pose), incre))
            self.last_imu = T_back

    def _save_keyframe(self, corner_full, surf_full):
        self.kf_corner.append(transform_cloud(corner_full, self.pose))
        self.kf_surf.append(transform_cloud(surf_full, self.pose))
        while len(self.kf_surf) >= WINDOW + 1:
            self.kf_surf.pop(0); self.kf_corner.pop(0)
        self.key_pose = self.pose.copy()
        self.keyframe_id += 1

    def push(self, pts, ring, initial_pose=None, info=None):
        """One LiDAR frame.  Returns the pose estimate [roll, pitch, yaw, x, y, z].  With `info` (cloud_info hints) the
        loop must run with its clamps disabled (rot_tolerance = z_tolerance = 0 in lm_params): transformUpdate (:976-1006)
        - IMU slerp, then the clamps - is applied here through `transform_update`."""
        self._update_initial_guess(initial_pose, info)
        f = self.be.extract_features(pts, ring, self.fprm) if self.fprm is not None else self.be.extract_features(pts, ring)
        ext = np.ascontiguousarray(pts[f["src_index"]], np.float32)
        corner_full = np.ascontiguousarray(ext[f["corner_idx"]]); surf_full = np.ascontiguousarray(ext[f["surf_idx"]])
        if self.first:
            self._save_keyframe(corner_full, surf_full)
            self.first = False
            self.trajectory.append(self.pose.copy()); self.results.append(None)
            return self.pose.copy()
        map_corner = self.be.voxel_grid(np.concatenate(self.kf_corner[::-1]), self.corner_leaf)
        map_surf = self.be.voxel_grid(np.concatenate(self.kf_surf[::-1]), self.surf_leaf)
        mid = self.be.map_create(map_corner, map_surf)
        corner = self.be.voxel_grid(corner_full, self.corner_leaf)
        surf = self.be.voxel_grid(surf_full, self.surf_leaf)
        pose, res = self.be.scan2map(mid, corner, surf, self.pose, self.prm)
        self.be.map_destroy(mid)
        if res.status != 1:                                   # "Not enough features": pose left at the prediction (:623-625)
            self.pose = np.asarray(pose, np.float32).copy()
            if self.transform_update is not None:
                imu_av = info is not None and bool(info["imu_available"])
                rp = info["imu_rpy"] if info is not None else (0.0, 0.0, 0.0)
                self.pose = self.transform_update(self.pose, imu_av, rp[0], rp[1], self.imu_rpy_weight)
            if not (res.deltaR == 100.0 and res.deltaT == 100.0):
                self.deltaR, self.deltaT = f32(res.deltaR), f32(res.deltaT)
        if float(self.deltaR) < 0.005 or float(self.deltaT) < 0.05:
            inc = _euler(_mul(_inv(_T16(self.key_pose)), _T16(self.pose)))     # calculateTranslation (:284-295)
            if self.keyframe_id <= 5 or abs(inc[2]) >= KEYFRAME_MIN_YAW or abs(inc[3]) >= KEYFRAME_MIN_DISTANCE or abs(inc[4]) >= KEYFRAME_MIN_DISTANCE:
                self._save_keyframe(corner_full, surf_full)
        self.trajectory.append(self.pose.copy()); self.results.append(res)
        return self.pose.copy()


class EngineBackend:
    """Adapter: lis_slam_b200.engine.Engine -> the backend protocol above."""

    def __init__(self, eng, gate_hint=1.0):
        self.eng, self.gate = eng, gate_hint

    def extract_features(self, pts, ring, prm=None):
        return self.eng.extract_features(pts, ring, prm)

    def voxel_grid(self, pts, leaf):
        return self.eng.voxel_grid(pts, leaf)

    def map_create(self, corner, surf):
        return self.eng.map_create(corner, surf, gate_hint=self.gate)

    def map_destroy(self, mid):
        self.eng.map_destroy(mid)

    def scan2map(self, mid, corner, surf, pose, prm):
        p, r, _ = self.eng.scan2map(mid, corner, surf, pose, prm)
        return p, r
