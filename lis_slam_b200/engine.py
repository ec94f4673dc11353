"""ctypes binding of liblisreg.so (include/lisreg.h) used by tests/ and bench.py.

This is plumbing only: every compute call goes through the C-ABI into the hand-written
sm_100a kernels.  There is no CPU fallback — if the shared library is missing or no CUDA
device is present the calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblisreg.so")
LUT_SIZE = 64
MAX_ITERS = 32

OK, NOT_ENOUGH_FEATURES, FEW_CORRESPONDENCES = 0, 1, 2


class LisregError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("stream", C.c_void_p), ("max_grid_cells", C.c_int32), ("own_stream", C.c_int32),
                ("reserved", C.c_int32 * 4)]


class LmParams(C.Structure):
    _fields_ = [
        ("max_iters", C.c_int32), ("early_exit", C.c_int32), ("sqdist_gate", C.c_float),
        ("conv_rot_deg", C.c_float), ("conv_trans_cm", C.c_float),
        ("edge_min_valid", C.c_int32), ("surf_min_valid", C.c_int32), ("min_sel", C.c_int32),
        ("degenerate_eig", C.c_float), ("use_label_weight", C.c_int32),
        ("label_score", C.c_float * LUT_SIZE), ("degenerate_in", C.c_int32),
        ("rot_tolerance", C.c_float), ("z_tolerance", C.c_float), ("want_iter_log", C.c_int32),
    ]


class LmIter(C.Structure):
    _fields_ = [
        ("AtA", C.c_float * 36), ("AtB", C.c_float * 6), ("X", C.c_float * 6), ("pose", C.c_float * 6),
        ("n_sel", C.c_int32), ("n_corner_sel", C.c_int32), ("n_surf_sel", C.c_int32), ("solved", C.c_int32),
        ("deltaR", C.c_float), ("deltaT", C.c_float),
    ]


class LmResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("iters", C.c_int32), ("converged", C.c_int32), ("is_degenerate", C.c_int32),
        ("n_sel_last", C.c_int32), ("deltaR", C.c_float), ("deltaT", C.c_float), ("pose", C.c_float * 6),
        ("n_corner", C.c_int32), ("n_surf", C.c_int32),
    ]


class BatchItem(C.Structure):
    _fields_ = [
        ("corner", C.c_void_p), ("clabel", C.c_void_p), ("surf", C.c_void_p), ("slabel", C.c_void_p),
        ("nc", C.c_int32), ("ns", C.c_int32), ("map_id", C.c_int32), ("reserved", C.c_int32),
    ]


class CloudLayout(C.Structure):
    _fields_ = [("point_step", C.c_int32), ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32),
                ("off_intensity", C.c_int32), ("off_ring", C.c_int32), ("off_time", C.c_int32), ("reserved", C.c_int32)]


LAYOUT_FLOAT4_RING, LAYOUT_XYZ_RING, LAYOUT_PCL_XYZIRT, LAYOUT_XYZ_SYNTH_RING = 0, 1, 2, 3


def cloud_layout(which):
    l = CloudLayout()
    lib().lisreg_cloud_layout_preset(C.byref(l), which)
    return l


class FeatParams(C.Structure):
    _fields_ = [("n_scan", C.c_int32), ("horizon", C.c_int32), ("downsample_rate", C.c_int32),
                ("min_range", C.c_float), ("max_range", C.c_float), ("edge_thr", C.c_float), ("surf_thr", C.c_float),
                ("reserved", C.c_int32), ("layout", CloudLayout)]


class FeatOut(C.Structure):
    _fields_ = [("n_extracted", C.c_int32), ("n_corner", C.c_int32), ("n_sharp", C.c_int32), ("n_flat", C.c_int32), ("n_surf", C.c_int32),
                ("src_index", C.c_void_p), ("col_ind", C.c_void_p), ("range", C.c_void_p),
                ("start_ring", C.c_void_p), ("end_ring", C.c_void_p),
                ("corner_idx", C.c_void_p), ("sharp_idx", C.c_void_p), ("flat_idx", C.c_void_p), ("surf_idx", C.c_void_p),
                ("curvature", C.c_void_p), ("label", C.c_void_p)]


class FrameParams(C.Structure):
    _fields_ = [("feat", FeatParams), ("corner_leaf", C.c_float), ("surf_leaf", C.c_float), ("lm", LmParams), ("deskew", C.c_void_p)]


class FrameItem(C.Structure):
    _fields_ = [("pts", C.c_void_p), ("ring", C.c_void_p), ("n", C.c_int32), ("map_id", C.c_int32)]


class OdomParams(C.Structure):
    _fields_ = [("frame", FrameParams), ("keyframe_min_distance", C.c_float), ("keyframe_min_yaw", C.c_float),
                ("window", C.c_int32), ("use_graph", C.c_int32), ("use_imu_heading_initialization", C.c_int32),
                ("imu_rpy_weight", C.c_float)]


class CloudInfo(C.Structure):
    """lisreg_cloud_info: the scalar hints of lis_slam::cloud_info (msg/cloud_info.msg:4-19)."""
    _fields_ = [("imu_available", C.c_int32), ("odom_available", C.c_int32), ("imu_roll_init", C.c_float),
                ("imu_pitch_init", C.c_float), ("imu_yaw_init", C.c_float), ("initial_guess", C.c_float * 6)]


class OdomResult(C.Structure):
    _fields_ = [("lm", LmResult), ("frame_id", C.c_int32), ("keyframe_id", C.c_int32), ("keyframe_saved", C.c_int32),
                ("map_rebuilt", C.c_int32), ("n_map_corner", C.c_int32), ("n_map_surf", C.c_int32), ("guess", C.c_float * 6)]


class SubmapInsertParams(C.Structure):
    _fields_ = [("dynamic_removal_on", C.c_int32), ("max_num_pts", C.c_int32), ("center_radius", C.c_float), ("dist_min", C.c_float),
                ("dist_max", C.c_float), ("near_dist", C.c_float)]


class SubmapInfo(C.Structure):
    _fields_ = [("n", C.c_int32 * 5), ("feature_point_num", C.c_int32), ("bound_min", C.c_double * 3), ("bound_max", C.c_double * 3),
                ("n_map_corner", C.c_int32), ("n_map_surf", C.c_int32)]


class LoopCandidate(C.Structure):
    _fields_ = [("submap_id", C.c_int32), ("use_epsc_init", C.c_int32), ("prekey_pose6", C.c_float * 6), ("epsc_T", C.c_float * 16),
                ("submap_pose6", C.c_float * 6)]


class LoopVerifyResult(C.Structure):
    _fields_ = [("found", C.c_int32), ("best", C.c_int32), ("best_score", C.c_double), ("correction", C.c_float * 16),
                ("key2pre", C.c_float * 16), ("t_correct", C.c_float * 16), ("constraint6", C.c_float * 6)]


class EpscCloud(C.Structure):
    _fields_ = [("corner", C.c_void_p), ("surf", C.c_void_p), ("sem", C.c_void_p), ("sem_label", C.c_void_p),
                ("nc", C.c_int32), ("ns", C.c_int32), ("nsem", C.c_int32), ("reserved", C.c_int32)]


class Deskew(C.Structure):
    _fields_ = [("imu_time", C.c_void_p), ("imu_rot", C.c_void_p), ("n_imu", C.c_int32), ("reserved", C.c_int32),
                ("time_scan_cur", C.c_double)]


class LoopParams(C.Structure):
    _fields_ = [("use_epsc", C.c_int32), ("use_sepsc", C.c_int32), ("use_fepsc", C.c_int32), ("use_pose", C.c_int32),
                ("skip_neighbour_distance", C.c_float), ("inflation_covariance", C.c_float), ("distance_threshold", C.c_float),
                ("reserved", C.c_int32)]


class LoopMatch(C.Structure):
    _fields_ = [("kind", C.c_int32), ("frame_id", C.c_int32), ("score", C.c_double), ("T", C.c_float * 16)]


class LoopResult(C.Structure):
    _fields_ = [("current_frame_id", C.c_int32), ("n_candidates", C.c_int32), ("n_matched", C.c_int32), ("reserved", C.c_int32),
                ("match", LoopMatch * 4)]


LOOP_KINDS = ("epsc", "sepsc", "fepsc", "pose")


class IcpParams(C.Structure):
    _fields_ = [("max_corr_dist", C.c_float), ("max_iters", C.c_int32), ("trans_eps", C.c_double), ("fitness_eps", C.c_double)]


class IcpPair(C.Structure):
    _fields_ = [("src", C.c_void_p), ("ns", C.c_int32), ("target_id", C.c_int32)]


class IcpResult(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("fitness", C.c_double), ("converged", C.c_int32), ("iters", C.c_int32),
                ("n_corr_last", C.c_int32), ("reserved", C.c_int32)]


class Profile(C.Structure):
    _fields_ = [
        ("lm_iter_ms", C.c_double), ("lm_iter_launches", C.c_int64), ("lm_alg_bytes", C.c_double),
        ("feat_ms", C.c_double), ("feat_launches", C.c_int64), ("feat_alg_bytes", C.c_double),
        ("voxel_ms", C.c_double), ("voxel_launches", C.c_int64), ("voxel_alg_bytes", C.c_double),
        ("index_ms", C.c_double), ("index_launches", C.c_int64), ("index_alg_bytes", C.c_double),
    ]


def build(force=False):
    """Compile liblisreg.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "lisreg.h"))
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", src_dir, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return LIB_PATH


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise LisregError("liblisreg.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                              "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, fp = C.c_void_p, C.c_int32, C.POINTER(C.c_float)
        L.lisreg_create.restype = i32
        L.lisreg_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.lisreg_destroy.argtypes = [vp]
        L.lisreg_last_error.restype = C.c_char_p
        L.lisreg_last_error.argtypes = [vp]
        L.lisreg_version.restype = C.c_char_p
        L.lisreg_sync.restype = i32
        L.lisreg_sync.argtypes = [vp]
        L.lisreg_launch_count.restype = C.c_int64
        L.lisreg_launch_count.argtypes = [vp]
        L.lisreg_lm_params_preset.argtypes = [C.POINTER(LmParams), C.c_char]
        L.lisreg_map_create.restype = i32
        L.lisreg_map_create.argtypes = [vp, vp, i32, vp, i32, C.c_float, C.POINTER(i32)]
        L.lisreg_map_create_dev.restype = i32
        L.lisreg_map_create_dev.argtypes = [vp, vp, i32, vp, i32, C.c_float, C.POINTER(i32)]
        L.lisreg_map_destroy.restype = i32
        L.lisreg_map_destroy.argtypes = [vp, i32]
        L.lisreg_knn5.restype = i32
        L.lisreg_knn5.argtypes = [vp, i32, i32, vp, i32, C.c_float, vp, vp]
        L.lisreg_scan2map.restype = i32
        L.lisreg_scan2map.argtypes = [vp, i32, vp, vp, i32, vp, vp, i32, fp, C.POINTER(LmParams), C.POINTER(LmResult), C.POINTER(LmIter)]
        L.lisreg_scan2map_batch.restype = i32
        L.lisreg_scan2map_batch.argtypes = [vp, i32, C.POINTER(BatchItem), fp, C.POINTER(LmParams), C.POINTER(LmResult), C.POINTER(LmIter)]
        L.lisreg_scan2map_batch_dev.restype = i32
        L.lisreg_scan2map_batch_dev.argtypes = [vp, i32, C.POINTER(BatchItem), vp, C.POINTER(LmParams), vp]
        L.lisreg_scan2map_batch_arena.restype = i32
        L.lisreg_scan2map_batch_arena.argtypes = [vp, i32, C.POINTER(BatchItem), vp, C.c_uint64, fp, C.POINTER(LmParams), C.POINTER(LmResult)]
        L.lisreg_feat_params_default.argtypes = [C.POINTER(FeatParams)]
        L.lisreg_cloud_layout_preset.argtypes = [C.POINTER(CloudLayout), i32]
        L.lisreg_extract_features.restype = i32
        L.lisreg_extract_features.argtypes = [vp, vp, vp, i32, C.POINTER(FeatParams), C.POINTER(FeatOut)]
        L.lisreg_voxel_grid.restype = i32
        L.lisreg_voxel_grid.argtypes = [vp, vp, i32, C.c_float, vp, C.POINTER(i32)]
        L.lisreg_frame_params_default.argtypes = [C.POINTER(FrameParams)]
        L.lisreg_frames_batch_dev.restype = i32
        L.lisreg_frames_batch_dev.argtypes = [vp, i32, C.POINTER(FrameItem), vp, C.POINTER(FrameParams), vp]
        L.lisreg_frames_batch_arena.restype = i32
        L.lisreg_frames_batch_arena.argtypes = [vp, i32, C.POINTER(FrameItem), vp, C.c_uint64, fp, C.POINTER(FrameParams), C.POINTER(LmResult)]
        L.lisreg_frames_batch_submit.restype = i32
        L.lisreg_frames_batch_submit.argtypes = [vp, i32, C.POINTER(FrameItem), vp, C.c_uint64, fp, C.POINTER(FrameParams), C.POINTER(i32)]
        L.lisreg_frames_batch_wait.restype = i32
        L.lisreg_frames_batch_wait.argtypes = [vp, i32, fp, C.POINTER(LmResult)]
        L.lisreg_epsc_describe.restype = i32
        L.lisreg_epsc_describe.argtypes = [vp, i32, C.POINTER(EpscCloud), vp, vp, vp, vp]
        L.lisreg_epsc_score_all.restype = i32
        L.lisreg_epsc_score_all.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        L.lisreg_epsc_score_all_dev.restype = i32
        L.lisreg_epsc_score_all_dev.argtypes = [vp, vp, i32, i32, vp, vp, vp]
        L.lisreg_icp_params_default.argtypes = [C.POINTER(IcpParams)]
        L.lisreg_map_distance_filter.restype = i32
        L.lisreg_map_distance_filter.argtypes = [vp, i32, i32, vp, i32, C.c_float, C.c_float, C.c_float, C.c_float, vp, C.POINTER(i32)]
        L.lisreg_extract_features_deskew.restype = i32
        L.lisreg_extract_features_deskew.argtypes = [vp, vp, vp, vp, i32, C.POINTER(FeatParams), C.POINTER(Deskew), C.POINTER(FeatOut), vp]
        L.lisreg_epsc_score_rows.restype = i32
        L.lisreg_epsc_score_rows.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp]
        L.lisreg_loop_params_default.argtypes = [C.POINTER(LoopParams)]
        L.lisreg_loop_create.restype = i32
        L.lisreg_loop_create.argtypes = [vp, C.POINTER(LoopParams), C.POINTER(C.c_uint8), C.POINTER(i32)]
        L.lisreg_loop_destroy.restype = i32
        L.lisreg_loop_destroy.argtypes = [vp, i32]
        L.lisreg_loop_detect.restype = i32
        L.lisreg_loop_detect.argtypes = [vp, i32, fp, i32, fp, i32, fp, C.POINTER(C.c_uint16), i32, fp, C.POINTER(LoopResult)]
        L.lisreg_icp_verify_batch.restype = i32
        L.lisreg_icp_verify_batch.argtypes = [vp, i32, C.POINTER(IcpPair), C.POINTER(IcpParams), C.POINTER(IcpResult)]
        L.lisreg_odom_params_default.argtypes = [C.POINTER(OdomParams)]
        L.lisreg_odom_create.restype = i32
        L.lisreg_odom_create.argtypes = [vp, C.POINTER(OdomParams), C.POINTER(i32)]
        L.lisreg_odom_destroy.restype = i32
        L.lisreg_odom_destroy.argtypes = [vp, i32]
        L.lisreg_odom_push.restype = i32
        L.lisreg_odom_push.argtypes = [vp, i32, vp, vp, i32, vp, fp, C.POINTER(OdomResult)]
        L.lisreg_odom_push_info.restype = i32
        L.lisreg_odom_push_info.argtypes = [vp, i32, vp, vp, i32, i32, C.POINTER(CloudInfo), fp, C.POINTER(OdomResult)]
        L.lisreg_transform_update.restype = None
        L.lisreg_transform_update.argtypes = [C.POINTER(CloudInfo), C.c_float, C.c_float, C.c_float, fp]
        L.lisreg_odom_push_dev.restype = i32
        L.lisreg_odom_push_dev.argtypes = [vp, i32, vp, vp, i32, vp, fp, C.POINTER(OdomResult)]
        L.lisreg_submap_create.restype = i32
        L.lisreg_submap_create.argtypes = [vp, C.POINTER(i32)]
        L.lisreg_submap_destroy.restype = i32
        L.lisreg_submap_destroy.argtypes = [vp, i32]
        L.lisreg_submap_clear.restype = i32
        L.lisreg_submap_clear.argtypes = [vp, i32]
        L.lisreg_submap_insert.restype = i32
        L.lisreg_submap_insert.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i32), fp, C.POINTER(SubmapInsertParams), C.POINTER(SubmapInfo)]
        L.lisreg_submap_extract.restype = i32
        L.lisreg_submap_extract.argtypes = [vp, i32, fp, fp, C.c_float, C.POINTER(i32), C.POINTER(SubmapInfo)]
        L.lisreg_submap_download.restype = i32
        L.lisreg_submap_download.argtypes = [vp, i32, i32, vp, i32, C.POINTER(i32)]
        L.lisreg_pretreat.restype = i32
        L.lisreg_pretreat.argtypes = [vp, vp, i32, i32, C.c_double, C.c_float, C.c_float, vp, vp, vp, C.POINTER(i32)]
        L.lisreg_deskew_constant_velocity.restype = i32
        L.lisreg_deskew_constant_velocity.argtypes = [vp, vp, vp, i32, C.c_float, fp, fp, vp]
        L.lisreg_loop_verify.restype = i32
        L.lisreg_loop_verify.argtypes = [vp, vp, i32, fp, fp, i32, C.POINTER(LoopCandidate), C.c_float, C.POINTER(IcpParams),
                                         C.POINTER(LoopVerifyResult), C.POINTER(IcpResult)]
        L.lisreg_selftest_smallmat.restype = i32
        L.lisreg_selftest_smallmat.argtypes = [vp, fp, fp, fp]
        L.lisreg_selftest_alu_peak.restype = i32
        L.lisreg_selftest_alu_peak.argtypes = [vp, C.POINTER(C.c_double)]
        for name in ("lisreg_comm_unique_id",):
            getattr(L, name).restype = i32
        L.lisreg_comm_unique_id.argtypes = [vp]
        L.lisreg_comm_init.restype = i32
        L.lisreg_comm_init.argtypes = [vp, i32, i32, vp]
        L.lisreg_comm_destroy.restype = i32
        L.lisreg_comm_destroy.argtypes = [vp]
        L.lisreg_allgather_results.restype = i32
        L.lisreg_allgather_results.argtypes = [vp, vp, vp, C.c_uint64]
        L.lisreg_allgather_wait.restype = i32
        L.lisreg_allgather_wait.argtypes = [vp, i32]
        L.lisreg_profile_enable.restype = i32
        L.lisreg_profile_enable.argtypes = [vp, i32]
        L.lisreg_profile_get.restype = i32
        L.lisreg_profile_get.argtypes = [vp, C.POINTER(Profile), i32]
        _LIB = L
    return _LIB


def lm_params(variant="A", **kw):
    p = LmParams()
    lib().lisreg_lm_params_preset(C.byref(p), variant.encode())
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def feat_params(**kw):
    p = FeatParams()
    lib().lisreg_feat_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def frame_params(variant="A", **lm_kw):
    p = FrameParams()
    lib().lisreg_frame_params_default(C.byref(p))
    lib().lisreg_lm_params_preset(C.byref(p.lm), variant.encode())
    for k, v in lm_kw.items():
        setattr(p.lm, k, v)
    return p


def cloud_info(imu_available=False, odom_available=False, imu_rpy=(0.0, 0.0, 0.0), initial_guess=(0, 0, 0, 0, 0, 0)):
    """lisreg_cloud_info from the message's scalar fields; initial_guess = (x, y, z, roll, pitch, yaw)."""
    ci = CloudInfo()
    ci.imu_available, ci.odom_available = int(bool(imu_available)), int(bool(odom_available))
    ci.imu_roll_init, ci.imu_pitch_init, ci.imu_yaw_init = (float(v) for v in imu_rpy)
    for i, v in enumerate(initial_guess):
        ci.initial_guess[i] = float(v)
    return ci


def transform_update(pose6, info, imu_rpy_weight=0.01, rot_tolerance=0.0, z_tolerance=0.0):
    """lisreg_transform_update: host arithmetic only (no device).  Returns the new pose6."""
    p = np.array(pose6, dtype=np.float32).copy()
    lib().lisreg_transform_update(None if info is None else C.byref(info), imu_rpy_weight, rot_tolerance, z_tolerance,
                                  p.ctypes.data_as(C.POINTER(C.c_float)))
    return p


def odom_params(variant="A", n_scan=64, use_graph=1, **lm_kw):
    p = OdomParams()
    lib().lisreg_odom_params_default(C.byref(p))
    lib().lisreg_lm_params_preset(C.byref(p.frame.lm), variant.encode())
    p.frame.feat.n_scan = n_scan
    p.use_graph = use_graph
    for k, v in lm_kw.items():
        setattr(p.frame.lm, k, v)
    return p


def _f4(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, "point clouds are (N,4) float32 {x,y,z,intensity}"
    return a


def _u16(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint16)


def _ptr(a):
    return None if a is None else a.ctypes.data


class Engine:
    """One lisreg context (one GPU, one stream)."""

    def __init__(self, device=0, stream=None, max_grid_cells=0, own_stream=False):
        self._h = C.c_void_p()
        cfg = Config(device=device, stream=stream, max_grid_cells=max_grid_cells, own_stream=1 if own_stream else 0)
        rc = lib().lisreg_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise LisregError("lisreg_create failed (%d): no CUDA device / driver — this engine has no CPU path" % rc)

    def close(self):
        if self._h:
            lib().lisreg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise LisregError("lisreg error %d: %s" % (rc, lib().lisreg_last_error(self._h).decode()))
        return rc

    def sync(self):
        self._ck(lib().lisreg_sync(self._h))

    @property
    def launches(self):
        return int(lib().lisreg_launch_count(self._h))

    # ---- maps
    def map_create(self, corner, surf, gate_hint=1.0):
        c, s = _f4(corner), _f4(surf)
        mid = C.c_int32(-1)
        self._ck(lib().lisreg_map_create(self._h, _ptr(c), len(c), _ptr(s), len(s), gate_hint, C.byref(mid)))
        return mid.value

    def map_create_dev(self, d_corner_ptr, mc, d_surf_ptr, ms, gate_hint=1.0):
        mid = C.c_int32(-1)
        self._ck(lib().lisreg_map_create_dev(self._h, d_corner_ptr, mc, d_surf_ptr, ms, gate_hint, C.byref(mid)))
        return mid.value

    def map_destroy(self, map_id):
        self._ck(lib().lisreg_map_destroy(self._h, map_id))

    def knn5(self, map_id, which, queries, gate):
        q = _f4(queries)
        idx = np.empty((len(q), 5), np.int32); sqd = np.empty((len(q), 5), np.float32)
        self._ck(lib().lisreg_knn5(self._h, map_id, which, _ptr(q), len(q), gate, _ptr(idx), _ptr(sqd)))
        return idx, sqd

    # ---- registration
    def scan2map(self, map_id, corner, surf, pose6, params, clabel=None, slabel=None, log=False):
        poses, res, logs = self.scan2map_batch([(map_id, corner, surf, clabel, slabel)], [pose6], params, log=log)
        return poses[0], res[0], (logs[0] if log else [])

    def scan2map_batch(self, regs, poses, params, log=False):
        """regs: list of (map_id, corner, surf, clabel|None, slabel|None). Host buffers; H2D, solve
        and D2H happen inside the call (this is the e2e path)."""
        B = len(regs)
        items = (BatchItem * B)()
        keep = []
        for b, (mid, c, s, cl, sl) in enumerate(regs):
            c, s, cl, sl = _f4(c), _f4(s), _u16(cl), _u16(sl)
            keep.append((c, s, cl, sl))
            items[b] = BatchItem(_ptr(c), _ptr(cl), _ptr(s), _ptr(sl), len(c), len(s), mid, 0)
        pose = np.ascontiguousarray(np.asarray(poses, dtype=np.float32).reshape(B, 6)).copy()
        res = (LmResult * B)()
        logs = None
        params.want_iter_log = 1 if log else 0
        if log:
            logs = (LmIter * (B * params.max_iters))()
        rc = self._ck(lib().lisreg_scan2map_batch(self._h, B, items, pose.ctypes.data_as(C.POINTER(C.c_float)),
                                                  C.byref(params), res, logs))
        out_logs = None
        if log:
            out_logs = [[logs[b * params.max_iters + i] for i in range(res[b].iters)] for b in range(B)]
        self.last_status = rc
        return pose, list(res), out_logs

    def scan2map_batch_arena(self, items, B, arena_ptr, arena_bytes, pose, params, res):
        """Frame packets packed in one (pinned) host arena; items hold byte offsets. pose: (B,6) f32
        numpy array in/out, res: ctypes array of LmResult."""
        return self._ck(lib().lisreg_scan2map_batch_arena(self._h, B, items, arena_ptr, arena_bytes,
                                                          pose.ctypes.data_as(C.POINTER(C.c_float)), C.byref(params), res))

    def extract_features(self, pts, ring, prm=None, time=None, imu_time=None, imu_rot=None, time_scan_cur=0.0):
        """F1-F5 on one raw sweep (host buffers). Returns a dict shaped like oracle.orc.extract_features.
        With time + IMU rotation table: motion de-skew (lisreg_extract_features_deskew); adds 'ext_pts' (M,4)."""
        prm = prm or feat_params()
        if prm.layout.point_step:      # caller-defined records (lisreg_cloud_layout): any contiguous array of n * point_step bytes
            p = np.ascontiguousarray(pts)
            assert p.nbytes % prm.layout.point_step == 0
            n_pts = p.nbytes // prm.layout.point_step
        else:
            p = _f4(pts); n_pts = len(p)
        r = None if ring is None else np.ascontiguousarray(ring, dtype=np.uint16)
        deskew = time is not None
        cap = prm.n_scan * prm.horizon
        a = {"src_index": np.zeros(cap, np.int32), "col_ind": np.zeros(cap, np.int32), "range": np.zeros(cap, np.float32),
             "start_ring": np.zeros(prm.n_scan, np.int32), "end_ring": np.zeros(prm.n_scan, np.int32),
             "corner_idx": np.zeros(prm.n_scan * 120, np.int32), "sharp_idx": np.zeros(prm.n_scan * 24, np.int32),
             "flat_idx": np.zeros(prm.n_scan * 60, np.int32), "surf_idx": np.zeros(cap, np.int32),
             "curvature": np.zeros(cap, np.float32), "label": np.zeros(cap, np.int32)}
        out = FeatOut()
        for k, v in a.items():
            setattr(out, k, v.ctypes.data)
        ext = None
        if deskew:
            t = np.ascontiguousarray(time, np.float32)
            it = np.ascontiguousarray(imu_time if imu_time is not None else np.zeros(0), np.float64)
            ir = np.ascontiguousarray(imu_rot if imu_rot is not None else np.zeros((0, 3)), np.float64).reshape(-1)
            dsk = Deskew(it.ctypes.data, ir.ctypes.data, len(it), 0, float(time_scan_cur))
            ext = np.zeros((cap, 4), np.float32)
            self._ck(lib().lisreg_extract_features_deskew(self._h, p.ctypes.data, _ptr(r), t.ctypes.data, n_pts, C.byref(prm),
                                                          C.byref(dsk), C.byref(out), ext.ctypes.data))
        else:
            self._ck(lib().lisreg_extract_features(self._h, p.ctypes.data, _ptr(r), n_pts, C.byref(prm), C.byref(out)))
        M = out.n_extracted
        res = {"M": M, "start_ring": a["start_ring"], "end_ring": a["end_ring"]}
        if ext is not None:
            res["ext_pts"] = ext[:M].copy()
        for k in ("src_index", "col_ind", "range", "curvature", "label"):
            res[k] = a[k][:M]
        res["corner_idx"] = a["corner_idx"][:out.n_corner]; res["sharp_idx"] = a["sharp_idx"][:out.n_sharp]
        res["flat_idx"] = a["flat_idx"][:out.n_flat]; res["surf_idx"] = a["surf_idx"][:out.n_surf]
        return res

    def pretreat(self, pts, n_scan, scan_period=0.1, min_range=0.0, max_range=70.0):
        """Ring / time synthesis (lisreg_pretreat). Returns (pts (m,4), ring (m,) u16, time (m,) f32)."""
        p = _f4(pts); n = len(p)
        out = np.zeros((n, 4), np.float32); ring = np.zeros(n, np.uint16); t = np.zeros(n, np.float32); m = C.c_int32(0)
        self._ck(lib().lisreg_pretreat(self._h, p.ctypes.data, n, n_scan, scan_period, min_range, max_range, out.ctypes.data, ring.ctypes.data,
                                       t.ctypes.data, C.byref(m)))
        return out[:m.value].copy(), ring[:m.value].copy(), t[:m.value].copy()

    def deskew_cv(self, pts, time, scan_period, lin_vel, ang_vel):
        p = _f4(pts); t = np.ascontiguousarray(time, np.float32)
        lv = np.ascontiguousarray(lin_vel, np.float32); av = np.ascontiguousarray(ang_vel, np.float32)
        out = np.zeros((max(len(p) - 1, 0), 4), np.float32)
        f = C.POINTER(C.c_float)
        self._ck(lib().lisreg_deskew_constant_velocity(self._h, p.ctypes.data, t.ctypes.data, len(p), scan_period, lv.ctypes.data_as(f),
                                                       av.ctypes.data_as(f), out.ctypes.data))
        return out

    def voxel_grid(self, pts, leaf):
        p = _f4(pts)
        out = np.zeros((max(len(p), 1), 4), np.float32)
        m = C.c_int32(0)
        self._ck(lib().lisreg_voxel_grid(self._h, p.ctypes.data, len(p), leaf, out.ctypes.data, C.byref(m)))
        return out[:m.value].copy()

    def frames_batch(self, frames, poses, params):
        """Whole-frame pipeline on host buffers. frames: list of (map_id, pts (n,4) f32, ring (n,) u16).
        Packs the sweeps into one arena (the e2e path). Returns (poses (F,6), [LmResult])."""
        F = len(frames)
        chunks, items, off = [], (FrameItem * F)(), 0
        for i, (mid, pts, ring) in enumerate(frames):
            p = _f4(pts); r = np.ascontiguousarray(ring, dtype=np.uint16)
            op = off; chunks.append(p.view(np.uint8).reshape(-1)); off += p.nbytes
            orr = off; chunks.append(r.view(np.uint8).reshape(-1)); off += r.nbytes
            pad = (-off) % 16
            if pad:
                chunks.append(np.zeros(pad, np.uint8)); off += pad
            items[i] = FrameItem(op, orr, len(p), mid)
        arena = np.concatenate(chunks) if chunks else np.zeros(16, np.uint8)
        pose = np.ascontiguousarray(np.asarray(poses, dtype=np.float32).reshape(F, 6)).copy()
        res = (LmResult * F)()
        self.last_status = self._ck(lib().lisreg_frames_batch_arena(self._h, F, items, arena.ctypes.data, arena.nbytes,
                                                                    pose.ctypes.data_as(C.POINTER(C.c_float)), C.byref(params), res))
        return pose, list(res)

    def frames_batch_arena(self, items, F, arena_ptr, arena_bytes, pose, params, res):
        return self._ck(lib().lisreg_frames_batch_arena(self._h, F, items, arena_ptr, arena_bytes,
                                                        pose.ctypes.data_as(C.POINTER(C.c_float)), C.byref(params), res))

    def frames_batch_submit(self, items, F, arena_ptr, arena_bytes, pose, params):
        """Asynchronous arena call: returns a ticket (two may be in flight); see lisreg_frames_batch_submit."""
        t = C.c_int32(-1)
        self._ck(lib().lisreg_frames_batch_submit(self._h, F, items, arena_ptr, arena_bytes,
                                                  pose.ctypes.data_as(C.POINTER(C.c_float)), C.byref(params), C.byref(t)))
        return t.value

    def frames_batch_wait(self, ticket, pose, res):
        return self._ck(lib().lisreg_frames_batch_wait(self._h, ticket, pose.ctypes.data_as(C.POINTER(C.c_float)), res))

    def frames_batch_dev(self, items, F, d_pose_ptr, params, d_res_ptr):
        return self._ck(lib().lisreg_frames_batch_dev(self._h, F, items, d_pose_ptr, C.byref(params), d_res_ptr))

    # ---- streaming odometry (device-resident sliding-window map) ----
    def odom_create(self, prm=None):
        prm = prm or odom_params()
        oid = C.c_int32(-1)
        self._ck(lib().lisreg_odom_create(self._h, C.byref(prm), C.byref(oid)))
        return oid.value

    def odom_destroy(self, oid):
        self._ck(lib().lisreg_odom_destroy(self._h, oid))

    def odom_push(self, oid, pts, ring, init_pose=None):
        """One sweep (host arrays) -> (pose6, OdomResult)."""
        p = _f4(pts); g = _u16(ring)
        pose = np.zeros(6, np.float32); res = OdomResult()
        ip = None if init_pose is None else np.ascontiguousarray(init_pose, np.float32)
        self.last_status = self._ck(lib().lisreg_odom_push(self._h, oid, p.ctypes.data, g.ctypes.data, len(p), _ptr(ip),
                                                           pose.ctypes.data_as(C.POINTER(C.c_float)), C.byref(res)))
        return pose, res

    def odom_push_info(self, oid, pts, ring, info, on_device=False, n=None):
        """One sweep with the cloud_info hints (CloudInfo or None).  on_device: pts / ring are device pointers, n given."""
        pose = np.zeros(6, np.float32); res = OdomResult()
        if on_device:
            pp, gp, cnt = pts, ring, n
        else:
            p = _f4(pts); g = _u16(ring); pp, gp, cnt = p.ctypes.data, g.ctypes.data, len(p)
        self.last_status = self._ck(lib().lisreg_odom_push_info(self._h, oid, pp, gp, cnt, 1 if on_device else 0,
                                                                None if info is None else C.byref(info),
                                                                pose.ctypes.data_as(C.POINTER(C.c_float)), C.byref(res)))
        return pose, res

    def odom_push_dev(self, oid, d_pts_ptr, d_ring_ptr, n, init_pose=None):
        pose = np.zeros(6, np.float32); res = OdomResult()
        ip = None if init_pose is None else np.ascontiguousarray(init_pose, np.float32)
        self.last_status = self._ck(lib().lisreg_odom_push_dev(self._h, oid, d_pts_ptr, d_ring_ptr, n, _ptr(ip),
                                                               pose.ctypes.data_as(C.POINTER(C.c_float)), C.byref(res)))
        return pose, res

    # ---- device-resident local map / submap ----
    def submap_create(self):
        sid = C.c_int32(-1)
        self._ck(lib().lisreg_submap_create(self._h, C.byref(sid)))
        return sid.value

    def submap_destroy(self, sid):
        self._ck(lib().lisreg_submap_destroy(self._h, sid))

    def submap_clear(self, sid):
        self._ck(lib().lisreg_submap_clear(self._h, sid))

    def submap_insert(self, sid, clouds5, pose6, dynrem=None, max_num_pts=20000):
        """clouds5: five (n,4) arrays (dynamic, pole, ground, building, outlier); dynrem = None or
        (center_radius, dist_min, dist_max, near).  Returns SubmapInfo."""
        arrs = [_f4(np.asarray(c, np.float32).reshape(-1, 4)) for c in clouds5]
        ptrs = (C.c_void_p * 5)(*[a.ctypes.data for a in arrs]); n = (C.c_int32 * 5)(*[len(a) for a in arrs])
        pose = np.ascontiguousarray(pose6, np.float32); info = SubmapInfo()
        prm = None
        if dynrem:
            prm = SubmapInsertParams(1, max_num_pts, dynrem[0], dynrem[1], dynrem[2], dynrem[3])
        self._ck(lib().lisreg_submap_insert(self._h, sid, ptrs, n, pose.ctypes.data_as(C.POINTER(C.c_float)),
                                            C.byref(prm) if prm else None, C.byref(info)))
        return info

    def submap_extract(self, sid, cur_pose6, leaf=None, gate_hint=2.0, map_id=-1):
        pose = np.ascontiguousarray(cur_pose6, np.float32); info = SubmapInfo(); mid = C.c_int32(map_id)
        lf = None if leaf is None else np.ascontiguousarray(leaf, np.float32)
        self._ck(lib().lisreg_submap_extract(self._h, sid, pose.ctypes.data_as(C.POINTER(C.c_float)),
                                             None if lf is None else lf.ctypes.data_as(C.POINTER(C.c_float)), gate_hint, C.byref(mid), C.byref(info)))
        return mid.value, info

    def submap_download(self, sid, cls):
        n = C.c_int32(0)
        self._ck(lib().lisreg_submap_download(self._h, sid, cls, None, 0, C.byref(n)))
        out = np.zeros((n.value, 4), np.float32)
        if n.value:
            self._ck(lib().lisreg_submap_download(self._h, sid, cls, out.ctypes.data, n.value, C.byref(n)))
        return out

    def loop_verify(self, key_cloud, key_pose6, key_rel_pose6, cands, fitness_threshold=0.5, prm=None):
        """detectLoopClosureForSubMap on device-resident submaps. cands: list of dicts {submap_id, use_epsc, prekey_pose6,
        epsc_T (4,4), submap_pose6}.  Returns (LoopVerifyResult, [IcpResult per candidate])."""
        k = _f4(key_cloud); P = len(cands)
        arr = (LoopCandidate * max(P, 1))()
        for i, c in enumerate(cands):
            arr[i].submap_id = c["submap_id"]; arr[i].use_epsc_init = 1 if c["use_epsc"] else 0
            arr[i].prekey_pose6 = (C.c_float * 6)(*[float(v) for v in c["prekey_pose6"]])
            arr[i].epsc_T = (C.c_float * 16)(*[float(v) for v in np.asarray(c["epsc_T"], np.float32).reshape(16)])
            arr[i].submap_pose6 = (C.c_float * 6)(*[float(v) for v in c["submap_pose6"]])
        if prm is None:
            prm = IcpParams(); lib().lisreg_icp_params_default(C.byref(prm))
        out = LoopVerifyResult(); per = (IcpResult * max(P, 1))()
        f = C.POINTER(C.c_float)
        a = np.ascontiguousarray(key_pose6, np.float32); b = np.ascontiguousarray(key_rel_pose6, np.float32)
        self._ck(lib().lisreg_loop_verify(self._h, k.ctypes.data, len(k), a.ctypes.data_as(f), b.ctypes.data_as(f), P, arr, fitness_threshold,
                                          C.byref(prm), C.byref(out), per))
        return out, list(per)[:P]

    def epsc_describe(self, clouds, using_map):
        """clouds: list of (corner (n,4), surf (n,4), sem (n,4), sem_label (n,)). Returns dict of (n,20,80) u8 arrays."""
        n = len(clouds)
        arr = (EpscCloud * n)(); keep = []
        for i, (c, s, m, l) in enumerate(clouds):
            c, s, m = _f4(c), _f4(s), _f4(m); l = np.ascontiguousarray(l, np.uint16)
            keep.append((c, s, m, l))
            arr[i] = EpscCloud(c.ctypes.data, s.ctypes.data, m.ctypes.data, l.ctypes.data, len(c), len(s), len(m), 0)
        lut = np.ascontiguousarray(using_map, np.uint8)
        out = [np.zeros((n, 20, 80), np.uint8) for _ in range(3)]
        self._ck(lib().lisreg_epsc_describe(self._h, n, arr, lut.ctypes.data, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data))
        return {"epsc": out[0], "sepsc": out[1], "fepsc": out[2]}

    def epsc_score_rows(self, desc, row_begin, row_stride, topk=5):
        """One rank's cyclic shard of the all-pairs scoring: rows row_begin, row_begin + row_stride, ..."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 1600)
        N = len(d)
        n_rows = max(0, (N - row_begin + row_stride - 1) // row_stride)
        idx = np.zeros((n_rows, topk), np.int32); score = np.zeros((n_rows, topk), np.float32); shift = np.zeros((n_rows, topk), np.int8)
        self._ck(lib().lisreg_epsc_score_rows(self._h, d.ctypes.data, N, row_begin, row_stride, topk, idx.ctypes.data, score.ctypes.data, shift.ctypes.data))
        return idx, score, shift

    def epsc_score_all(self, desc, topk=5):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 1600)
        N = len(d)
        idx = np.zeros((N, topk), np.int32); score = np.zeros((N, topk), np.float32); shift = np.zeros((N, topk), np.int8)
        self._ck(lib().lisreg_epsc_score_all(self._h, d.ctypes.data, N, topk, idx.ctypes.data, score.ctypes.data, shift.ctypes.data))
        return idx, score, shift

    def target_create(self, pts):
        """Registers an ICP target cloud (stored as the 'surf' cloud of a map slot)."""
        return self.map_create(np.zeros((0, 4), np.float32), pts, gate_hint=1.0)

    def map_distance_filter(self, map_id, which, feat, center_radius=30.0, dyn_min=0.3, dyn_max=3.0, near=0.03):
        """map_scan_feature_pts_distance_removal (subMap.h:1063-1098). Returns the keep mask (n,) bool."""
        f = _f4(feat)
        keep = np.zeros(max(len(f), 1), np.uint8)
        nk = C.c_int32(0)
        self._ck(lib().lisreg_map_distance_filter(self._h, map_id, which, f.ctypes.data, len(f), center_radius, dyn_min, dyn_max, near,
                                                  keep.ctypes.data, C.byref(nk)))
        return keep[:len(f)].astype(bool)

    def loop_create(self, lut, use_epsc=False, use_sepsc=False, use_fepsc=True, use_pose=False):
        """EPSCGeneration instance (loop detector) on the device; returns its id."""
        prm = LoopParams()
        lib().lisreg_loop_params_default(C.byref(prm))
        prm.use_epsc, prm.use_sepsc, prm.use_fepsc, prm.use_pose = int(use_epsc), int(use_sepsc), int(use_fepsc), int(use_pose)
        lut = np.ascontiguousarray(lut, np.uint8)
        det = C.c_int32(-1)
        self._ck(lib().lisreg_loop_create(self._h, C.byref(prm), lut.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(det)))
        return det.value

    def loop_destroy(self, det):
        self._ck(lib().lisreg_loop_destroy(self._h, det))

    def loop_detect(self, det, corner, surf, sem, sem_label, odom):
        """EPSCGeneration::loopDetection. Returns (current_frame_id, n_candidates, [(kind, matched_id, score, T 4x4)])."""
        c, s, m = _f4(corner), _f4(surf), _f4(sem)
        lab = np.ascontiguousarray(sem_label, np.uint16)
        od = np.ascontiguousarray(odom, np.float32).reshape(16)
        res = LoopResult()
        fpt = C.POINTER(C.c_float)
        self._ck(lib().lisreg_loop_detect(self._h, det, c.ctypes.data_as(fpt), len(c), s.ctypes.data_as(fpt), len(s), m.ctypes.data_as(fpt),
                                          lab.ctypes.data_as(C.POINTER(C.c_uint16)), len(m), od.ctypes.data_as(fpt), C.byref(res)))
        out = [(LOOP_KINDS[res.match[i].kind], res.match[i].frame_id, res.match[i].score,
                np.array(res.match[i].T, np.float32).reshape(4, 4)) for i in range(res.n_matched)]
        return res.current_frame_id, res.n_candidates, out

    def icp_verify_batch(self, pairs, prm=None):
        """pairs: list of (src (n,4) already pre-transformed by the initial guess, target_id). Returns [IcpResult]."""
        P = len(pairs)
        if prm is None:
            prm = IcpParams(); lib().lisreg_icp_params_default(C.byref(prm))
        arr = (IcpPair * P)(); keep = []
        for i, (src, tid) in enumerate(pairs):
            s = _f4(src); keep.append(s)
            arr[i] = IcpPair(s.ctypes.data, len(s), tid)
        out = (IcpResult * P)()
        self._ck(lib().lisreg_icp_verify_batch(self._h, P, arr, C.byref(prm), out))
        return list(out)

    def selftest_smallmat(self, A, b):
        A = np.ascontiguousarray(A, np.float32).reshape(36); b = np.ascontiguousarray(b, np.float32).reshape(6)
        out = np.zeros(98, np.float32)
        self._ck(lib().lisreg_selftest_smallmat(self._h, A.ctypes.data_as(C.POINTER(C.c_float)), b.ctypes.data_as(C.POINTER(C.c_float)),
                                                out.ctypes.data_as(C.POINTER(C.c_float))))
        return {"E": out[:6], "V": out[6:42].reshape(6, 6), "X": out[42:48], "qr_ok": int(out[48]),
                "inv": out[49:85].reshape(6, 6), "lu_ok": int(out[85]), "W3": out[86:89], "V3": out[89:98].reshape(3, 3)}

    def alu_peak(self):
        v = C.c_double(0)
        self._ck(lib().lisreg_selftest_alu_peak(self._h, C.byref(v)))
        return v.value

    # ---- multi-GPU exchange (NCCL behind the C-ABI) ----
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        if lib().lisreg_comm_unique_id(buf) != 0:
            raise LisregError("lisreg_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def comm_init(self, world, rank, uid):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(lib().lisreg_comm_init(self._h, world, rank, buf))

    def comm_destroy(self):
        self._ck(lib().lisreg_comm_destroy(self._h))

    def allgather_results(self, d_send_ptr, d_recv_ptr, bytes_per_rank):
        self._ck(lib().lisreg_allgather_results(self._h, d_send_ptr, d_recv_ptr, bytes_per_rank))

    def allgather_wait(self, back=0):
        self._ck(lib().lisreg_allgather_wait(self._h, back))

    def profile_enable(self, on=True):
        self._ck(lib().lisreg_profile_enable(self._h, 1 if on else 0))

    def profile_get(self, reset=True):
        p = Profile()
        self._ck(lib().lisreg_profile_get(self._h, C.byref(p), 1 if reset else 0))
        return p

    def scan2map_batch_dev(self, items, B, d_pose_ptr, params, d_res_ptr):
        """All buffers resident in HBM (items = ctypes array of BatchItem holding DEVICE pointers).
        Asynchronous on the context stream."""
        return self._ck(lib().lisreg_scan2map_batch_dev(self._h, B, items, d_pose_ptr, C.byref(params), d_res_ptr))
