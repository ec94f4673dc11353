"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liblisreg_oracle.so.

Imported only by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline and
--impl reference).  The product package lis_slam_b200 never imports this module.
PARITY UNPINNED by the reference (no tests/golden vectors upstream); see DESIGN.md.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
LUT_SIZE = 64


def build(force=False):
    so = os.path.join(_HERE, "liblisreg_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.startswith("orc_") and f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return so


class LmParams(C.Structure):
    _fields_ = [
        ("max_iters", C.c_int32), ("early_exit", C.c_int32), ("sqdist_gate", C.c_float),
        ("conv_rot_deg", C.c_float), ("conv_trans_cm", C.c_float),
        ("edge_min_valid", C.c_int32), ("surf_min_valid", C.c_int32), ("min_sel", C.c_int32),
        ("degenerate_eig", C.c_float), ("use_label_weight", C.c_int32),
        ("label_score", C.c_float * LUT_SIZE), ("degenerate_in", C.c_int32),
        ("rot_tolerance", C.c_float), ("z_tolerance", C.c_float), ("n_threads", C.c_int32),
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
        ("n_sel_last", C.c_int32), ("deltaR", C.c_float), ("deltaT", C.c_float),
        ("ms_build", C.c_double), ("ms_iters", C.c_double),
    ]


class FeatParams(C.Structure):
    _fields_ = [("n_scan", C.c_int32), ("horizon", C.c_int32), ("downsample_rate", C.c_int32),
                ("min_range", C.c_float), ("max_range", C.c_float), ("edge_thr", C.c_float), ("surf_thr", C.c_float)]


def feat_params(n_scan=64, horizon=1800, downsample_rate=1, min_range=0.0, max_range=70.0, edge_thr=1.0, surf_thr=0.1):
    return FeatParams(n_scan, horizon, downsample_rate, min_range, max_range, edge_thr, surf_thr)


class IcpParams(C.Structure):
    _fields_ = [("max_corr_dist", C.c_float), ("max_iters", C.c_int32), ("trans_eps", C.c_double), ("fitness_eps", C.c_double)]


class IcpResult(C.Structure):
    _fields_ = [("T", C.c_float * 16), ("fitness", C.c_double), ("converged", C.c_int32), ("iters", C.c_int32),
                ("n_corr_last", C.c_int32), ("pad", C.c_int32)]


def icp_params(max_corr_dist=10.0, max_iters=30, trans_eps=1e-4, fitness_eps=1e-4):
    return IcpParams(max_corr_dist, max_iters, trans_eps, fitness_eps)


# label_sorce of config/label.yaml:214-234 (reference values)
LABEL_SCORE = [1.0, 1.0, 0.6, 0.5, 0.8, 0.5, 0.5, 0.5, 0.5, 1.2, 1.2, 1.2, 0.5, 1.0, 0.8, 0.5, 1.3, 0.5, 1.5, 1.5]


def lm_params(variant="A", **kw):
    """Parameter presets of the three copies of the loop."""
    p = LmParams()
    p.early_exit = 1
    p.edge_min_valid, p.surf_min_valid, p.min_sel = -1, 100, 50
    p.degenerate_eig = 100.0
    p.rot_tolerance = p.z_tolerance = 1000.0
    p.n_threads = 1
    for i, v in enumerate(LABEL_SCORE):
        p.label_score[i] = v
    if variant == "A":
        p.max_iters, p.sqdist_gate, p.conv_rot_deg, p.conv_trans_cm, p.use_label_weight = 15, 1.0, 0.005, 0.05, 0
    elif variant == "B":
        p.max_iters, p.sqdist_gate, p.conv_rot_deg, p.conv_trans_cm, p.use_label_weight = 20, 2.0, 0.003, 0.03, 1
    elif variant == "C":
        p.max_iters, p.sqdist_gate, p.conv_rot_deg, p.conv_trans_cm, p.use_label_weight = 30, 2.0, 0.002, 0.02, 1
    else:
        raise ValueError(variant)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp, ip, u16p = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint16)
        L.orc_scan2map.restype = C.c_int
        L.orc_scan2map.argtypes = [fp, u16p, C.c_int32, fp, u16p, C.c_int32, fp, C.c_int32, fp, C.c_int32,
                                   fp, C.POINTER(LmParams), C.POINTER(LmResult), C.POINTER(LmIter)]
        L.orc_kdtree_build.restype = C.c_void_p
        L.orc_kdtree_build.argtypes = [fp, C.c_int32]
        L.orc_kdtree_free.argtypes = [C.c_void_p]
        L.orc_knn_batch.argtypes = [C.c_void_p, fp, C.c_int32, C.c_int32, ip, fp, C.c_int32]
        L.orc_corner_coeff.restype = C.c_int
        L.orc_corner_coeff.argtypes = [fp, fp, fp]
        L.orc_surf_coeff.restype = C.c_int
        L.orc_surf_coeff.argtypes = [fp, fp, fp]
        L.orc_pose_to_affine.argtypes = [fp, fp]
        L.orc_jacobi_eigen_f32.argtypes = [fp, C.c_int32, fp, fp]
        L.orc_qr_solve_f32.restype = C.c_int
        L.orc_qr_solve_f32.argtypes = [fp, C.c_int32, fp, fp]
        L.orc_lu_inv_f32.restype = C.c_int
        L.orc_lu_inv_f32.argtypes = [fp, C.c_int32, fp]
        L.orc_plane_fit_5x3.argtypes = [fp, fp]
        L.orc_project_scan.restype = C.c_int32
        L.orc_project_scan.argtypes = [fp, u16p, C.c_int32, C.POINTER(FeatParams), ip, ip, fp, ip, ip]
        L.orc_extract_features.argtypes = [fp, ip, C.c_int32, ip, ip, C.POINTER(FeatParams), ip, ip, ip, ip, ip, ip, ip, ip, fp, ip]
        L.orc_voxel_grid.restype = C.c_int32
        L.orc_voxel_grid.argtypes = [fp, C.c_int32, C.c_float, fp, C.c_int32]
        u8p = C.POINTER(C.c_uint8)
        L.orc_epsc_describe.argtypes = [fp, C.c_int32, fp, C.c_int32, fp, u16p, C.c_int32, u8p, u8p, u8p, u8p]
        L.orc_epsc_distance.restype = C.c_double
        L.orc_epsc_distance.argtypes = [u8p, u8p, ip, ip]
        L.orc_epsc_score_all.argtypes = [u8p, C.c_int32, C.c_int32, ip, fp, C.POINTER(C.c_int8), C.c_int32]
        L.orc_icp.restype = C.c_int
        L.orc_icp.argtypes = [fp, C.c_int32, fp, C.c_int32, C.POINTER(IcpParams), C.POINTER(IcpResult)]
        L.orc_deskew.argtypes = [fp, fp, ip, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32, C.c_double, fp]
        L.orc_map_distance_filter.restype = C.c_int32
        L.orc_map_distance_filter.argtypes = [fp, C.c_int32, fp, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, u8p]
        L.orc_loop_create.restype = C.c_void_p
        L.orc_loop_create.argtypes = [u8p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.orc_loop_free.argtypes = [C.c_void_p]
        L.orc_loop_detect.restype = C.c_int32
        L.orc_loop_detect.argtypes = [C.c_void_p, fp, C.c_int32, fp, C.c_int32, fp, u16p, C.c_int32, fp, ip, ip, ip, ip,
                                      C.POINTER(C.c_double), fp]
        L.orc_loop_project.argtypes = [fp, u16p, C.c_int32, fp]
        L.orc_loop_global_icp.argtypes = [fp, fp, C.c_float, fp]
        _LIB = L
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _u16(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.uint16)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint16))


def scan2map(corner, surf, map_corner, map_surf, pose6, params, clabel=None, slabel=None, log=True):
    """Returns (pose6_out, LmResult, [LmIter...])."""
    L = lib()
    c, cp = _f(corner); s, sp = _f(surf); mc, mcp = _f(map_corner); ms, msp = _f(map_surf)
    cl, clp = _u16(clabel); sl, slp = _u16(slabel)
    pose = np.array(pose6, dtype=np.float32).copy()
    res = LmResult()
    logs = (LmIter * params.max_iters)() if log else None
    L.orc_scan2map(cp, clp, len(c), sp, slp, len(s), mcp, len(mc), msp, len(ms),
                   pose.ctypes.data_as(C.POINTER(C.c_float)), C.byref(params), C.byref(res), logs)
    return pose, res, (list(logs)[: res.iters] if log else [])


def knn(map_pts4, queries4, k=5, n_threads=1):
    L = lib()
    m, mp = _f(map_pts4); q, qp = _f(queries4)
    t = L.orc_kdtree_build(mp, len(m))
    idx = np.empty((len(q), k), np.int32); sqd = np.empty((len(q), k), np.float32)
    L.orc_knn_batch(t, qp, len(q), k, idx.ctypes.data_as(C.POINTER(C.c_int32)), sqd.ctypes.data_as(C.POINTER(C.c_float)), n_threads)
    L.orc_kdtree_free(t)
    return idx, sqd


def jacobi_eigen(A):
    A, ap = _f(A); n = A.shape[0]
    W = np.empty(n, np.float32); V = np.empty((n, n), np.float32)
    lib().orc_jacobi_eigen_f32(ap, n, W.ctypes.data_as(C.POINTER(C.c_float)), V.ctypes.data_as(C.POINTER(C.c_float)))
    return W, V


def qr_solve(A, b):
    A, ap = _f(A); b, bp = _f(b); n = A.shape[0]
    x = np.empty(n, np.float32)
    ok = lib().orc_qr_solve_f32(ap, n, bp, x.ctypes.data_as(C.POINTER(C.c_float)))
    return ok, x


def lu_inv(A):
    A, ap = _f(A); n = A.shape[0]
    out = np.empty((n, n), np.float32)
    ok = lib().orc_lu_inv_f32(ap, n, out.ctypes.data_as(C.POINTER(C.c_float)))
    return ok, out


def plane_fit(A53):
    A, ap = _f(A53)
    x = np.empty(3, np.float32)
    lib().orc_plane_fit_5x3(ap, x.ctypes.data_as(C.POINTER(C.c_float)))
    return x


def corner_coeff(q, nb):
    q, qp = _f(q); nb, nbp = _f(nb)
    c = np.zeros(4, np.float32)
    ok = lib().orc_corner_coeff(qp, nbp, c.ctypes.data_as(C.POINTER(C.c_float)))
    return ok, c


def surf_coeff(q, nb):
    q, qp = _f(q); nb, nbp = _f(nb)
    c = np.zeros(4, np.float32)
    ok = lib().orc_surf_coeff(qp, nbp, c.ctypes.data_as(C.POINTER(C.c_float)))
    return ok, c


def pose_to_affine(pose6):
    p, pp = _f(pose6)
    T = np.empty(12, np.float32)
    lib().orc_pose_to_affine(pp, T.ctypes.data_as(C.POINTER(C.c_float)))
    return T.reshape(3, 4)


def extract_features(pts4, ring, prm=None):
    """F1-F5 on one raw sweep. Returns dict of arrays (index lists refer to the extracted cloud)."""
    L = lib()
    prm = prm or feat_params()
    p, pp = _f(pts4)
    r = np.ascontiguousarray(ring, dtype=np.uint16)
    cap = prm.n_scan * prm.horizon
    ip = C.POINTER(C.c_int32)
    src = np.zeros(cap, np.int32); col = np.zeros(cap, np.int32); rng = np.zeros(cap, np.float32)
    sr = np.zeros(prm.n_scan, np.int32); er = np.zeros(prm.n_scan, np.int32)
    M = L.orc_project_scan(pp, r.ctypes.data_as(C.POINTER(C.c_uint16)), len(p), C.byref(prm), src.ctypes.data_as(ip),
                           col.ctypes.data_as(ip), rng.ctypes.data_as(C.POINTER(C.c_float)), sr.ctypes.data_as(ip), er.ctypes.data_as(ip))
    corner = np.zeros(prm.n_scan * 120, np.int32); sharp = np.zeros(prm.n_scan * 24, np.int32)
    flat = np.zeros(prm.n_scan * 60, np.int32); surf = np.zeros(cap, np.int32)
    n = [C.c_int32(0) for _ in range(4)]
    curv = np.zeros(max(M, 1), np.float32); label = np.zeros(max(M, 1), np.int32)
    L.orc_extract_features(rng.ctypes.data_as(C.POINTER(C.c_float)), col.ctypes.data_as(ip), M, sr.ctypes.data_as(ip), er.ctypes.data_as(ip),
                           C.byref(prm), corner.ctypes.data_as(ip), C.byref(n[0]), sharp.ctypes.data_as(ip), C.byref(n[1]),
                           flat.ctypes.data_as(ip), C.byref(n[2]), surf.ctypes.data_as(ip), C.byref(n[3]),
                           curv.ctypes.data_as(C.POINTER(C.c_float)), label.ctypes.data_as(ip))
    return {"M": M, "src_index": src[:M], "col_ind": col[:M], "range": rng[:M], "start_ring": sr, "end_ring": er,
            "corner_idx": corner[:n[0].value], "sharp_idx": sharp[:n[1].value], "flat_idx": flat[:n[2].value],
            "surf_idx": surf[:n[3].value], "curvature": curv[:M], "label": label[:M]}


def voxel_grid(pts4, leaf):
    p, pp = _f(pts4)
    out = np.zeros((max(len(p), 1), 4), np.float32)
    m = lib().orc_voxel_grid(pp, len(p), leaf, out.ctypes.data_as(C.POINTER(C.c_float)), len(out))
    return out[:m].copy()


# using_label of config/label.yaml:187-206: label -> class {10 dynamic, 40 ground, 50 building, 81 pole, 70 outlier}
USING_LABEL = {1: 10, 2: 10, 3: 10, 4: 10, 5: 10, 6: 10, 7: 10, 8: 10, 9: 40, 10: 40, 11: 40, 12: 70, 13: 50, 14: 50,
               15: 70, 16: 81, 17: 70, 18: 81, 19: 81}


def using_map_lut():
    lut = np.zeros(256, np.uint8)
    for k, v in USING_LABEL.items():
        lut[k] = v
    return lut


def epsc_describe(corner4, surf4, sem4, sem_label, lut=None):
    c, cp = _f(corner4); s, sp = _f(surf4); m, mp = _f(sem4)
    lab = np.ascontiguousarray(sem_label, np.uint16)
    lut = using_map_lut() if lut is None else np.ascontiguousarray(lut, np.uint8)
    u8p = C.POINTER(C.c_uint8)
    out = [np.zeros(1600, np.uint8) for _ in range(3)]
    lib().orc_epsc_describe(cp, len(c), sp, len(s), mp, lab.ctypes.data_as(C.POINTER(C.c_uint16)), len(m),
                            lut.ctypes.data_as(u8p), *[o.ctypes.data_as(u8p) for o in out])
    return {"epsc": out[0].reshape(20, 80), "sepsc": out[1].reshape(20, 80), "fepsc": out[2].reshape(20, 80)}


def epsc_distance(d1, d2):
    a = np.ascontiguousarray(d1, np.uint8).reshape(-1); b = np.ascontiguousarray(d2, np.uint8).reshape(-1)
    u8p = C.POINTER(C.c_uint8)
    sh, sad = C.c_int32(0), C.c_int32(0)
    sc = lib().orc_epsc_distance(a.ctypes.data_as(u8p), b.ctypes.data_as(u8p), C.byref(sh), C.byref(sad))
    return sc, sh.value, sad.value


def epsc_score_all(desc, topk=5, n_threads=1):
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 1600)
    N = len(d)
    idx = np.zeros((N, topk), np.int32); score = np.zeros((N, topk), np.float32); shift = np.zeros((N, topk), np.int8)
    lib().orc_epsc_score_all(d.ctypes.data_as(C.POINTER(C.c_uint8)), N, topk, idx.ctypes.data_as(C.POINTER(C.c_int32)),
                             score.ctypes.data_as(C.POINTER(C.c_float)), shift.ctypes.data_as(C.POINTER(C.c_int8)), n_threads)
    return idx, score, shift


def icp(src4, tgt4, prm=None):
    s, sp = _f(src4); t, tp = _f(tgt4)
    prm = prm or icp_params()
    res = IcpResult()
    lib().orc_icp(sp, len(s), tp, len(t), C.byref(prm), C.byref(res))
    return np.array(res.T, np.float32).reshape(4, 4), res



LOOP_KINDS = ("epsc", "sepsc", "fepsc", "pose")


class LoopDetector:
    """EPSCGeneration::loopDetection (epscGeneration.cpp:663-992): stateful, append-only."""

    def __init__(self, use_epsc=False, use_sepsc=False, use_fepsc=True, use_pose=False, lut=None):
        lut = using_map_lut() if lut is None else np.ascontiguousarray(lut, np.uint8)
        self._h = lib().orc_loop_create(lut.ctypes.data_as(C.POINTER(C.c_uint8)), int(use_epsc), int(use_sepsc), int(use_fepsc), int(use_pose))

    def detect(self, corner4, surf4, sem4, sem_label, odom):
        """Returns (current_frame_id, n_candidates, [(kind, matched_id, score, T 4x4)])."""
        c, cp = _f(corner4); s, sp = _f(surf4); m, mp = _f(sem4)
        lab = np.ascontiguousarray(sem_label, np.uint16)
        od = np.ascontiguousarray(odom, np.float32).reshape(16)
        cur, ncand = C.c_int32(0), C.c_int32(0)
        kinds = np.zeros(4, np.int32); ids = np.zeros(4, np.int32); scores = np.zeros(4, np.float64); T = np.zeros((4, 16), np.float32)
        ip = C.POINTER(C.c_int32)
        n = lib().orc_loop_detect(self._h, cp, len(c), sp, len(s), mp, lab.ctypes.data_as(C.POINTER(C.c_uint16)), len(m),
                                  od.ctypes.data_as(C.POINTER(C.c_float)), C.byref(cur), C.byref(ncand), kinds.ctypes.data_as(ip),
                                  ids.ctypes.data_as(ip), scores.ctypes.data_as(C.POINTER(C.c_double)), T.ctypes.data_as(C.POINTER(C.c_float)))
        return cur.value, ncand.value, [(LOOP_KINDS[kinds[i]], int(ids[i]), float(scores[i]), T[i].reshape(4, 4).copy()) for i in range(n)]

    def close(self):
        if self._h:
            lib().orc_loop_free(self._h); self._h = None


def loop_project(sem4, sem_label):
    m, mp = _f(sem4)
    lab = np.ascontiguousarray(sem_label, np.uint16)
    out = np.zeros((360, 4), np.float32)
    lib().orc_loop_project(mp, lab.ctypes.data_as(C.POINTER(C.c_uint16)), len(m), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def loop_global_icp(proj1, proj2, yaw_diff):
    a = np.ascontiguousarray(proj1, np.float32); b = np.ascontiguousarray(proj2, np.float32)
    T = np.zeros(16, np.float32)
    fp = C.POINTER(C.c_float)
    lib().orc_loop_global_icp(a.ctypes.data_as(fp), b.ctypes.data_as(fp), float(yaw_diff), T.ctypes.data_as(fp))
    return T.reshape(4, 4)


def deskew(pts4, time, src_index, imu_time, imu_rot, time_scan_cur):
    """deskewPoint over the extracted points (laserProcessing.cpp:427-462). imu_rot: (n,3) doubles. Returns (M,4)."""
    p, pp = _f(pts4)
    t = np.ascontiguousarray(time, np.float32)
    si = np.ascontiguousarray(src_index, np.int32)
    it = np.ascontiguousarray(imu_time, np.float64); ir = np.ascontiguousarray(imu_rot, np.float64).reshape(-1)
    out = np.zeros((len(si), 4), np.float32)
    dp = C.POINTER(C.c_double)
    lib().orc_deskew(pp, t.ctypes.data_as(C.POINTER(C.c_float)), si.ctypes.data_as(C.POINTER(C.c_int32)), len(si),
                     it.ctypes.data_as(dp), ir.ctypes.data_as(dp), len(it), float(time_scan_cur), out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def map_distance_filter(feat4, map4, center_radius=30.0, dyn_min=0.3, dyn_max=3.0, near=0.03):
    """map_scan_feature_pts_distance_removal (subMap.h:1063-1098). Returns the keep mask (n,) bool."""
    f, fp_ = _f(feat4); m, mp = _f(map4)
    keep = np.zeros(len(f), np.uint8)
    lib().orc_map_distance_filter(fp_, len(f), mp, len(m), center_radius, dyn_min, dyn_max, near, keep.ctypes.data_as(C.POINTER(C.c_uint8)))
    return keep.astype(bool)


class Submap:
    """localMap_t / submap_t class clouds + insert_local_map + extractSlidingCloud (oracle/orc_submap.cpp)."""
    LEAF = (0.1, 0.05, 0.4, 0.2, 0.6)          # dynamic, pole, ground, building, outlier (subMapOptmizationNode.cpp:1393-1397)

    def __init__(self):
        L = lib()
        L.orc_submap_create.restype = C.c_void_p
        L.orc_submap_free.argtypes = [C.c_void_p]; L.orc_submap_clear.argtypes = [C.c_void_p]
        L.orc_submap_get.restype = C.c_int32
        self.h = C.c_void_p(L.orc_submap_create())
        self.bound = np.zeros(6, np.float64)

    def close(self):
        if self.h:
            lib().orc_submap_free(self.h); self.h = None

    def clear(self):
        lib().orc_submap_clear(self.h)

    def insert(self, clouds5, pose6, dynrem=None, max_num_pts=20000):
        """clouds5: five (n,4) arrays; dynrem = None or (center_radius, dist_min, dist_max, near). Returns counts[5]."""
        arrs = [np.ascontiguousarray(c, np.float32).reshape(-1, 4) for c in clouds5]
        ptrs = (C.c_void_p * 5)(*[a.ctypes.data for a in arrs]); n = (C.c_int32 * 5)(*[len(a) for a in arrs])
        pose = np.ascontiguousarray(pose6, np.float32); counts = (C.c_int32 * 5)()
        dr = dynrem or (30.0, 0.3, 3.0, 0.03)
        lib().orc_submap_insert(self.h, ptrs, n, pose.ctypes.data_as(C.c_void_p), C.c_int32(1 if dynrem else 0), C.c_int32(max_num_pts),
                                C.c_float(dr[0]), C.c_float(dr[1]), C.c_float(dr[2]), C.c_float(dr[3]), counts, self.bound.ctypes.data_as(C.c_void_p))
        return list(counts)

    def extract(self, cur_pose6, leaf=None):
        leaf = np.ascontiguousarray(leaf or self.LEAF, np.float32)
        cap = sum(self.count(c) for c in range(5)) + 1
        corner = np.zeros((cap, 4), np.float32); surf = np.zeros((cap, 4), np.float32)
        nc, ns = C.c_int32(0), C.c_int32(0); counts = (C.c_int32 * 5)()
        pose = np.ascontiguousarray(cur_pose6, np.float32)
        lib().orc_submap_extract(self.h, pose.ctypes.data_as(C.c_void_p), leaf.ctypes.data_as(C.c_void_p), self.bound.ctypes.data_as(C.c_void_p),
                                 corner.ctypes.data_as(C.c_void_p), C.byref(nc), surf.ctypes.data_as(C.c_void_p), C.byref(ns), counts)
        return corner[:nc.value].copy(), surf[:ns.value].copy(), list(counts)

    def count(self, c):
        return lib().orc_submap_get(self.h, C.c_int32(c), None, C.c_int32(0))

    def get(self, c):
        n = self.count(c)
        out = np.zeros((n, 4), np.float32)
        lib().orc_submap_get(self.h, C.c_int32(c), out.ctypes.data_as(C.c_void_p), C.c_int32(n))
        return out


def pretreat(pts4, n_scan, scan_period=0.1, min_range=0.0, max_range=70.0):
    """Ring / time synthesis (laserPretreatmentNode.cpp:60-230). Returns (pts (m,4), ring (m,) u16, time (m,) f32)."""
    p, pp = _f(pts4)
    out = np.zeros((len(p), 4), np.float32); ring = np.zeros(len(p), np.uint16); t = np.zeros(len(p), np.float32)
    L = lib(); L.orc_pretreat.restype = C.c_int32
    m = L.orc_pretreat(pp, C.c_int32(len(p)), C.c_int32(n_scan), C.c_double(scan_period), C.c_float(min_range), C.c_float(max_range),
                       out.ctypes.data_as(C.c_void_p), ring.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
    return out[:m].copy(), ring[:m].copy(), t[:m].copy()


def deskew_cv(pts4, time, scan_period, lin_vel, ang_vel):
    """DistortionAdjust::AdjustCloud (distortionAdjust.cpp:419-479). Returns (n - 1, 4)."""
    p, pp = _f(pts4)
    t = np.ascontiguousarray(time, np.float32); lv = np.ascontiguousarray(lin_vel, np.float32); av = np.ascontiguousarray(ang_vel, np.float32)
    out = np.zeros((max(len(p) - 1, 0), 4), np.float32)
    L = lib(); L.orc_deskew_cv.restype = C.c_int32
    m = L.orc_deskew_cv(pp, t.ctypes.data_as(C.c_void_p), C.c_int32(len(p)), C.c_float(scan_period), lv.ctypes.data_as(C.c_void_p),
                        av.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out[:m]


def loop_verify(key_cloud, key_pose6, key_rel_pose6, cands, fitness_threshold=0.5, prm=None):
    """detectLoopClosureForSubMap. cands: list of dicts {submap: Submap, use_epsc, prekey_pose6, epsc_T (4,4), submap_pose6}.
    Returns dict(found, best, best_score, correction, key2pre, t_correct, constraint6, fitness[], converged[])."""
    k, kp = _f(key_cloud)
    P = len(cands)
    cp = np.zeros((max(P, 1), 29), np.float32)
    hs = (C.c_void_p * max(P, 1))()
    for i, c in enumerate(cands):
        cp[i, 0] = 1.0 if c["use_epsc"] else 0.0
        cp[i, 1:7] = c["prekey_pose6"]; cp[i, 7:23] = np.asarray(c["epsc_T"], np.float32).reshape(16); cp[i, 23:29] = c["submap_pose6"]
        hs[i] = c["submap"].h
    prm = prm or icp_params()
    best = C.c_int32(-1); score = C.c_double(0)
    corr = np.zeros(16, np.float32); k2p = np.zeros(16, np.float32); tc = np.zeros(16, np.float32); c6 = np.zeros(6, np.float32)
    fit = np.zeros(max(P, 1), np.float64); conv = np.zeros(max(P, 1), np.int32)
    a = np.ascontiguousarray(key_pose6, np.float32); b = np.ascontiguousarray(key_rel_pose6, np.float32)
    vp = C.c_void_p
    L = lib(); L.orc_loop_verify.restype = C.c_int32
    found = L.orc_loop_verify(kp, C.c_int32(len(k)), a.ctypes.data_as(vp), b.ctypes.data_as(vp), C.c_int32(P), hs, cp.ctypes.data_as(vp),
                              C.c_float(fitness_threshold), C.byref(prm), C.byref(best), C.byref(score), corr.ctypes.data_as(vp),
                              k2p.ctypes.data_as(vp), tc.ctypes.data_as(vp), c6.ctypes.data_as(vp), fit.ctypes.data_as(vp), conv.ctypes.data_as(vp))
    return dict(found=found, best=best.value, best_score=score.value, correction=corr.reshape(4, 4), key2pre=k2p.reshape(4, 4),
                t_correct=tc.reshape(4, 4), constraint6=c6, fitness=fit[:P], converged=conv[:P])


def transform_update(pose6, imu_available, imu_roll, imu_pitch, imu_rpy_weight, rot_tol=0.0, z_tol=0.0):
    """transformUpdate (odomEstimationNode.cpp:976-1006): IMU roll / pitch slerp + clamps.  Returns the new pose6."""
    L = lib()
    L.orc_transform_update.restype = None
    L.orc_transform_update.argtypes = [C.POINTER(C.c_float), C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]
    p = np.array(pose6, dtype=np.float32).copy()
    L.orc_transform_update(p.ctypes.data_as(C.POINTER(C.c_float)), int(bool(imu_available)), float(imu_roll), float(imu_pitch),
                           float(imu_rpy_weight), float(rot_tol), float(z_tol))
    return p
