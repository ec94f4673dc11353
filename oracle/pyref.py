"""ORACLE #2 — TEST INFRASTRUCTURE ONLY (never imported by the product; see oracle/orc_linalg.h for the rule).

An INDEPENDENT restatement of the scan-to-map loop in Python, written from the reference source
(src/node/odomEstimationNode.cpp:243-258, 596-1006; variants B/C src/node/subMapOptmizationNode.cpp:1509-2001,
4485-4976) and built on LIBRARY routines instead of hand-written ones, so that a misreading shared by the C++ oracle
(oracle/orc_*.cpp) and the CUDA kernels - both written by the same hand - cannot hide:

    nearest neighbours     scipy.spatial.cKDTree            (reference: pcl::KdTreeFLANN, exact L2 k-NN)
    3x3 / 6x6 eigen        cv2.eigen                        (the reference calls cv::eigen itself, :690, :928)
    6x6 solve              cv2.solve(..., DECOMP_QR)        (the reference's own call, :921)
    A^T A, A^T b, matP     cv2.gemm / cv2.transpose / cv2.invert   (cv::Mat operator*, .inv(), :918-920, :945)
    plane fit              Householder QR with column pivoting in fp32 numpy (Eigen colPivHouseholderQr, :783);
                           cross-checked against numpy.linalg.lstsq in the tests
    voxel grid             numpy (PCL VoxelGrid published algorithm: voxel_grid.hpp applyFilter)
    descriptor distance    numpy integer arithmetic (epscGeneration.cpp:633-660)

Per-point arithmetic is numpy float32 in the reference's expression order; the double-literal sub-expressions
(`cx + 0.1 * v`, `1 - 0.9 * fabs(d)`, `> 0.2`) are promoted to float64 exactly where C++ promotes them.
sin / cos follow the repo-wide resolution (DESIGN.md numerics): correctly rounded float of the double routine.
tests/test_pyref_oracle.py checks the C++ oracle against this file: pose and step <= 1e-6, selection counts equal,
A^T A <= 5e-6 relative (1e-6 typical), per iteration.  PARITY UNPINNED by the reference itself (it ships no vectors and cannot be
built here); this file pins the C++ oracle to an independent reading plus the reference's own library calls.
"""
import numpy as np

f32 = np.float32
f64 = np.float64


# ------------------------------------------------------------------------------------------------ parameters
LABEL_SCORE = [1.0, 1.0, 0.6, 0.5, 0.8, 0.5, 0.5, 0.5, 0.5, 1.2, 1.2, 1.2, 0.5, 1.0, 0.8, 0.5, 1.3, 0.5, 1.5, 1.5]   # config/label.yaml:214-234


def params(variant="A", **kw):
    """Constants hard-coded in the three copies of the loop."""
    p = dict(max_iters=15, gate=1.0, conv_rot=0.005, conv_trans=0.05, use_w=False,            # A: odomEstimationNode.cpp:606, :657, :965
             edge_min=-1, surf_min=100, min_sel=50, eig_thr=100.0, early_exit=True,
             rot_tol=1000.0, z_tol=1000.0)                                                   # config/params.yaml:123-124
    if variant == "B":
        p.update(max_iters=20, gate=2.0, conv_rot=0.003, conv_trans=0.03, use_w=True)        # subMapOptmizationNode.cpp:1520, :1610, :1963
    elif variant == "C":
        p.update(max_iters=30, gate=2.0, conv_rot=0.002, conv_trans=0.02, use_w=True)        # :4500, :4962
    p.update(kw)
    return p


# ------------------------------------------------------------------------------------------------ F7
def _sin(x):
    return f32(np.sin(f64(x)))


def _cos(x):
    return f32(np.cos(f64(x)))


def get_transformation(x, y, z, roll, pitch, yaw):
    """pcl::getTransformation (common/impl/eigen.hpp), Scalar = float: 3x4 float32."""
    A, B, C, D, E, F = _cos(yaw), _sin(yaw), _cos(pitch), _sin(pitch), _cos(roll), _sin(roll)
    DE, DF = f32(D * E), f32(D * F)
    T = np.zeros((3, 4), f32)
    T[0] = [A * C, A * DF - B * E, B * F + A * DE, x]
    T[1] = [B * C, A * E + B * DF, B * DE - A * F, y]
    T[2] = [-D, C * F, C * E, z]
    return T


def trans2affine(pose6):
    """trans2Affine3f (common.cpp:55-58): pose6 = [roll, pitch, yaw, x, y, z]."""
    t = np.asarray(pose6, f32)
    return get_transformation(t[3], t[4], t[5], t[0], t[1], t[2])


def associate(T, p):
    """pointAssociateToMap (:243-258): three products and three additions per coordinate, left to right, fp32."""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    q = np.empty((len(p), 3), f32)
    for r in range(3):
        q[:, r] = T[r, 0] * x + T[r, 1] * y + T[r, 2] * z + T[r, 3]
    return q


# ------------------------------------------------------------------------------------------------ F8
class Knn:
    """Exact 5-NN with squared distances in the FLANN L2 functor's fp32 operation order, ascending, ties by index.
    cKDTree works in float64; 8 candidates are fetched and re-ranked in fp32 so that a pair of neighbours whose
    distances differ only below fp32 resolution cannot change the set."""

    def __init__(self, cloud4):
        from scipy.spatial import cKDTree
        self.pts = np.ascontiguousarray(np.asarray(cloud4, f32)[:, :3])
        self.tree = cKDTree(self.pts.astype(f64)) if len(self.pts) else None

    def query5(self, q):
        n = len(q)
        idx = np.full((n, 5), -1, np.int64); sqd = np.full((n, 5), np.inf, f32)
        m = len(self.pts)
        if m == 0 or n == 0:
            return idx, sqd
        k = min(8, m)
        _, cand = self.tree.query(q.astype(f64), k=k)
        cand = cand.reshape(n, k)
        nb = self.pts[cand]                                   # n x k x 3
        dx = q[:, None, 0] - nb[:, :, 0]; dy = q[:, None, 1] - nb[:, :, 1]; dz = q[:, None, 2] - nb[:, :, 2]
        d = dx * dx; d = d + dy * dy; d = d + dz * dz         # fp32, FLANN order
        order = np.lexsort((cand, d), axis=1)                 # by d, then by index
        cand = np.take_along_axis(cand, order, 1); d = np.take_along_axis(d, order, 1)
        kk = min(5, k)
        idx[:, :kk] = cand[:, :kk]; sqd[:, :kk] = d[:, :kk]
        return idx, sqd


# ------------------------------------------------------------------------------------------------ F9
def corner_optimization(scan, T, knn, map4, gate, weights=None):
    """cornerOptimization (:633-747).  Returns (selected original points n x 3, coeff n x 4 = s*(la,lb,lc,ld2))."""
    import cv2
    if len(scan) == 0:
        return np.zeros((0, 3), f32), np.zeros((0, 4), f32)
    sel = associate(T, scan)
    idx, sqd = knn.query5(sel)
    ok5 = sqd[:, 4] < f32(gate)                                # pointSearchSqDis[4] < 1.0  (B/C: size()==5 && < 2.0)
    out_p, out_c = [], []
    mp = np.asarray(map4, f32)
    for i in np.nonzero(ok5)[0]:
        nb = mp[idx[i], :3]
        cx = cy = cz = f32(0)
        for j in range(5):
            cx = f32(cx + nb[j, 0]); cy = f32(cy + nb[j, 1]); cz = f32(cz + nb[j, 2])
        cx = f32(cx / f32(5)); cy = f32(cy / f32(5)); cz = f32(cz / f32(5))
        a11 = a12 = a13 = a22 = a23 = a33 = f32(0)
        for j in range(5):
            ax = f32(nb[j, 0] - cx); ay = f32(nb[j, 1] - cy); az = f32(nb[j, 2] - cz)
            a11 = f32(a11 + f32(ax * ax)); a12 = f32(a12 + f32(ax * ay)); a13 = f32(a13 + f32(ax * az))
            a22 = f32(a22 + f32(ay * ay)); a23 = f32(a23 + f32(ay * az)); a33 = f32(a33 + f32(az * az))
        a11, a12, a13, a22, a23, a33 = (f32(v / f32(5)) for v in (a11, a12, a13, a22, a23, a33))
        matA1 = np.array([[a11, a12, a13], [a12, a22, a23], [a13, a23, a33]], f32)
        _, matD1, matV1 = cv2.eigen(matA1)                      # eigenvalues descending, eigenvectors in rows
        matD1 = matD1.reshape(-1)
        if not (matD1[0] > f32(3) * matD1[1]):
            continue
        x0, y0, z0 = sel[i]
        x1 = f32(f64(cx) + 0.1 * f64(matV1[0, 0])); y1 = f32(f64(cy) + 0.1 * f64(matV1[0, 1])); z1 = f32(f64(cz) + 0.1 * f64(matV1[0, 2]))
        x2 = f32(f64(cx) - 0.1 * f64(matV1[0, 0])); y2 = f32(f64(cy) - 0.1 * f64(matV1[0, 1])); z2 = f32(f64(cz) - 0.1 * f64(matV1[0, 2]))
        m11 = f32(f32((x0 - x1) * (y0 - y2)) - f32((x0 - x2) * (y0 - y1)))
        m12 = f32(f32((x0 - x1) * (z0 - z2)) - f32((x0 - x2) * (z0 - z1)))
        m13 = f32(f32((y0 - y1) * (z0 - z2)) - f32((y0 - y2) * (z0 - z1)))
        with np.errstate(all="ignore"):
            a012 = np.sqrt(f32(f32(f32(m11 * m11) + f32(m12 * m12)) + f32(m13 * m13)))
            l12 = np.sqrt(f32(f32(f32((x1 - x2) * (x1 - x2)) + f32((y1 - y2) * (y1 - y2))) + f32((z1 - z2) * (z1 - z2))))
            la = f32(f32(f32(f32((y1 - y2) * m11) + f32((z1 - z2) * m12)) / a012) / l12)
            lb = f32(f32(-f32(f32((x1 - x2) * m11) - f32((z1 - z2) * m13)) / a012) / l12)
            lc = f32(f32(-f32(f32((x1 - x2) * m12) + f32((y1 - y2) * m13)) / a012) / l12)
            ld2 = f32(a012 / l12)
        s = f32(1.0 - 0.9 * f64(np.abs(ld2)))
        w = f32(1) if weights is None else weights[i]
        if f64(s) > 0.1:                                         # the gate stays on the UNWEIGHTED s (B: :1683)
            ws = f32(w * s) if weights is not None else s
            out_p.append(scan[i, :3]); out_c.append([f32(ws * la), f32(ws * lb), f32(ws * lc), f32(ws * ld2)])
    if not out_p:
        return np.zeros((0, 3), f32), np.zeros((0, 4), f32)
    return np.asarray(out_p, f32), np.asarray(out_c, f32)


# ------------------------------------------------------------------------------------------------ F10
def plane_fit_colpiv_qr(A0):
    """x = argmin |A0 x - (-1)| for a batch of 5x3 systems: Householder QR with column pivoting, fp32 throughout
    (Eigen::ColPivHouseholderQR<Matrix<float,5,3>>::solve).  A0: n x 5 x 3 float32 -> n x 3 float32."""
    n = len(A0)
    R = np.array(A0, f32, copy=True)
    b = np.full((n, 5), -1, f32)
    perm = np.tile(np.arange(3), (n, 1))
    rows = np.arange(n)
    with np.errstate(all="ignore"):
        for k in range(3):
            # pivot: remaining column with the largest squared norm over rows k..4
            norms = np.zeros((n, 3), f32)
            for c in range(k, 3):
                acc = np.zeros(n, f32)
                for r in range(k, 5):
                    acc = acc + R[:, r, c] * R[:, r, c]
                norms[:, c] = acc
            norms[:, :k] = -1
            piv = np.argmax(norms, axis=1)
            tmp = R[rows, :, k].copy(); R[rows, :, k] = R[rows, :, piv]; R[rows, :, piv] = tmp
            tp = perm[rows, k].copy(); perm[rows, k] = perm[rows, piv]; perm[rows, piv] = tp
            # Householder vector for column k (Eigen makeHouseholderInPlace): v = [1, essential], tau, beta
            c0 = R[:, k, k].copy()
            tail = np.zeros(n, f32)
            for r in range(k + 1, 5):
                tail = tail + R[:, r, k] * R[:, r, k]
            beta = np.sqrt(c0 * c0 + tail).astype(f32)
            beta = np.where(c0 >= 0, -beta, beta).astype(f32)
            degenerate = tail <= np.finfo(f32).tiny
            denom = (c0 - beta).astype(f32)
            ess = np.zeros((n, 5), f32)
            for r in range(k + 1, 5):
                ess[:, r] = np.where(degenerate, 0, R[:, r, k] / denom)
            tau = np.where(degenerate, 0, (beta - c0) / beta).astype(f32)
            beta = np.where(degenerate, c0, beta).astype(f32)
            # apply H = I - tau v v^T to the remaining columns and to b
            for c in range(k + 1, 3):
                dot = R[:, k, c].copy()
                for r in range(k + 1, 5):
                    dot = dot + ess[:, r] * R[:, r, c]
                R[:, k, c] = R[:, k, c] - tau * dot
                for r in range(k + 1, 5):
                    R[:, r, c] = R[:, r, c] - tau * dot * ess[:, r]
            dot = b[:, k].copy()
            for r in range(k + 1, 5):
                dot = dot + ess[:, r] * b[:, r]
            b[:, k] = b[:, k] - tau * dot
            for r in range(k + 1, 5):
                b[:, r] = b[:, r] - tau * dot * ess[:, r]
            R[:, k, k] = beta
            for r in range(k + 1, 5):
                R[:, r, k] = 0
        # back substitution on the 3x3 triangle, then undo the permutation
        y = np.zeros((n, 3), f32)
        for k in (2, 1, 0):
            acc = b[:, k].copy()
            for c in range(k + 1, 3):
                acc = acc - R[:, k, c] * y[:, c]
            y[:, k] = acc / R[:, k, k]
        x = np.zeros((n, 3), f32)
        x[rows[:, None], perm] = y
    return x


def surf_optimization(scan, T, knn, map4, gate, weights=None):
    """surfOptimization (:749-827).  Returns (selected original points, coeff = s*(pa,pb,pc,pd2))."""
    if len(scan) == 0:
        return np.zeros((0, 3), f32), np.zeros((0, 4), f32)
    sel = associate(T, scan)
    idx, sqd = knn.query5(sel)
    ok5 = np.nonzero(sqd[:, 4] < f32(gate))[0]
    if len(ok5) == 0:
        return np.zeros((0, 3), f32), np.zeros((0, 4), f32)
    mp = np.asarray(map4, f32)
    nb = mp[idx[ok5]][:, :, :3]                               # n x 5 x 3
    X = plane_fit_colpiv_qr(nb)
    with np.errstate(all="ignore"):
        pa, pb, pc = X[:, 0].copy(), X[:, 1].copy(), X[:, 2].copy()
        pd = np.ones(len(X), f32)
        ps = np.sqrt(pa * pa + pb * pb + pc * pc)
        pa = pa / ps; pb = pb / ps; pc = pc / ps; pd = pd / ps
        valid = np.ones(len(X), bool)
        for j in range(5):
            v = pa * nb[:, j, 0] + pb * nb[:, j, 1] + pc * nb[:, j, 2] + pd
            valid &= ~(np.abs(v).astype(f64) > 0.2)
        q = sel[ok5]
        pd2 = pa * q[:, 0] + pb * q[:, 1] + pc * q[:, 2] + pd
        rr = np.sqrt(np.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2]))
        s = (1.0 - 0.9 * np.abs(pd2).astype(f64) / rr.astype(f64)).astype(f32)
        keep = valid & (s.astype(f64) > 0.1)
    ws = s if weights is None else (weights[ok5] * s).astype(f32)
    coeff = np.stack([ws * pa, ws * pb, ws * pc, ws * pd2], 1).astype(f32)
    return scan[ok5][keep][:, :3].astype(f32), coeff[keep]


# ------------------------------------------------------------------------------------------------ F12
def lm_optimization(pose, ori, coeff, iter_count, state, prm):
    """LMOptimization (:852-974).  pose: float32[6] updated in place.  state: {'degenerate': bool} persists like the
    member isDegenerate.  Returns (converged, record)."""
    import cv2
    rec = dict(n_sel=len(ori), solved=False)
    srx, crx = _sin(pose[1]), _cos(pose[1])
    sry, cry = _sin(pose[2]), _cos(pose[2])
    srz, crz = _sin(pose[0]), _cos(pose[0])
    n = len(ori)
    if n < prm["min_sel"]:
        return False, rec
    # lidar -> camera
    px, py, pz = ori[:, 1], ori[:, 2], ori[:, 0]
    cx, cy, cz, ci = coeff[:, 1], coeff[:, 2], coeff[:, 0], coeff[:, 3]
    arx = (crx * sry * srz * px + crx * crz * sry * py - srx * sry * pz) * cx + \
          (-srx * srz * px - crz * srx * py - crx * pz) * cy + \
          (crx * cry * srz * px + crx * cry * crz * py - cry * srx * pz) * cz
    ary = ((cry * srx * srz - crz * sry) * px + (sry * srz + cry * crz * srx) * py + crx * cry * pz) * cx + \
          ((-cry * crz - srx * sry * srz) * px + (cry * srz - crz * srx * sry) * py - crx * sry * pz) * cz
    arz = ((crz * srx * sry - cry * srz) * px + (-cry * crz - srx * sry * srz) * py) * cx + \
          (crx * crz * px - crx * srz * py) * cy + \
          ((sry * srz + cry * crz * srx) * px + (crz * sry - cry * srx * srz) * py) * cz
    matA = np.ascontiguousarray(np.stack([arz, arx, ary, cz, cx, cy], 1), f32)
    matB = np.ascontiguousarray((-ci).reshape(-1, 1), f32)
    assert matA.dtype == f32
    matAt = cv2.transpose(matA)
    matAtA = cv2.gemm(matAt, matA, 1.0, None, 0.0)
    matAtB = cv2.gemm(matAt, matB, 1.0, None, 0.0)
    ok, matX = cv2.solve(matAtA, matAtB, flags=cv2.DECOMP_QR)
    if not ok:
        matX = np.zeros((6, 1), f32)
    matP = np.zeros((6, 6), f32)                                # the LOCAL matP (:880) shadows the member: zero unless iter 0 (Q1)
    if iter_count == 0:
        _, matE, matV = cv2.eigen(matAtA)
        matE = matE.reshape(-1)
        matV2 = matV.copy()
        state["degenerate"] = False
        for i in range(5, -1, -1):
            if matE[i] < f32(prm["eig_thr"]):
                matV2[i, :] = 0
                state["degenerate"] = True
            else:
                break
        _, Vinv = cv2.invert(matV)                              # cv::Mat::inv() default DECOMP_LU
        matP = cv2.gemm(Vinv, matV2, 1.0, None, 0.0)
    if state["degenerate"]:
        matX = cv2.gemm(matP, matX.copy(), 1.0, None, 0.0)
    X = matX.reshape(-1).astype(f32)
    for k in range(6):
        pose[k] = f32(pose[k] + X[k])
    r2d = f32(180.0 / np.pi)                                    # pcl::rad2deg(float): alpha * 57.29578f
    deltaR = f32(np.sqrt(sum(f64(f32(X[k] * r2d)) ** 2 for k in range(3))))
    deltaT = f32(np.sqrt(sum(f64(f32(X[k] * f32(100))) ** 2 for k in range(3, 6))))
    rec.update(solved=True, AtA=matAtA.copy(), AtB=matAtB.reshape(-1).copy(), X=X.copy(), deltaR=deltaR, deltaT=deltaT)
    return bool(f64(deltaR) < prm["conv_rot"] and f64(deltaT) < prm["conv_trans"]), rec


# ------------------------------------------------------------------------------------------------ F13
def scan2map(corner, surf, map_corner, map_surf, pose6, prm, clabel=None, slabel=None):
    """scan2SubMapOptimization (:596-626) + the clamps of transformUpdate (:1001-1003).
    Returns (pose float32[6], dict(status, iters, converged, degenerate), per-iteration records)."""
    pose = np.array(pose6, f32, copy=True)
    corner = np.asarray(corner, f32); surf = np.asarray(surf, f32)
    info = dict(status=0, iters=0, converged=False, degenerate=False)
    if not (len(corner) > prm["edge_min"] and len(surf) > prm["surf_min"]):
        info["status"] = 1
        return pose, info, []
    kc, ks = Knn(map_corner), Knn(map_surf)
    wc = ws = None
    if prm["use_w"]:
        lut = np.zeros(65536, f32); lut[:len(LABEL_SCORE)] = LABEL_SCORE
        wc = (2.0 - lut[np.asarray(clabel, np.int64)].astype(f64)).astype(f32) if clabel is not None else np.full(len(corner), 1.0, f32)
        ws = (2.0 - lut[np.asarray(slabel, np.int64)].astype(f64)).astype(f32) if slabel is not None else np.full(len(surf), 1.0, f32)
    state = {"degenerate": False}
    log = []
    any_small = False
    for it in range(prm["max_iters"]):
        T = trans2affine(pose)
        po_c, co_c = corner_optimization(corner, T, kc, map_corner, prm["gate"], wc)
        po_s, co_s = surf_optimization(surf, T, ks, map_surf, prm["gate"], ws)
        ori = np.concatenate([po_c, po_s]); coeff = np.concatenate([co_c, co_s])      # combineOptimizationCoeffs
        conv, rec = lm_optimization(pose, ori, coeff, it, state, prm)
        rec.update(n_corner_sel=len(po_c), n_surf_sel=len(po_s), pose=pose.copy())
        log.append(rec)
        info["iters"] = it + 1
        any_small |= not rec["solved"]
        if conv and prm["early_exit"]:
            info["converged"] = True
            break
        info["converged"] = conv
    if prm["rot_tol"] > 0:
        pose[0] = np.clip(pose[0], -f32(prm["rot_tol"]), f32(prm["rot_tol"]))
        pose[1] = np.clip(pose[1], -f32(prm["rot_tol"]), f32(prm["rot_tol"]))
    if prm["z_tol"] > 0:
        pose[5] = np.clip(pose[5], -f32(prm["z_tol"]), f32(prm["z_tol"]))
    info["degenerate"] = state["degenerate"]
    info["status"] = 2 if any_small else 0
    return pose, info, log


# ------------------------------------------------------------------------------------------------ F6
def voxel_grid(pts4, leaf):
    """pcl::VoxelGrid<PointXYZI>::applyFilter with downsample_all_data (default) and min_points_per_voxel 0.
    Output: one centroid (x, y, z, intensity; fp32 accumulation in input order) per occupied voxel, ascending voxel
    index.  (std::sort's order inside a voxel is implementation-defined upstream; input order is the resolution used
    everywhere in this repo.)"""
    p = np.asarray(pts4, f32)
    if len(p) == 0:
        return np.zeros((0, 4), f32)
    inv = f32(1.0) / f32(leaf)
    mn = p[:, :3].min(0); mx = p[:, :3].max(0)
    d = [np.int64(f32(f32(mx[k] - mn[k]) * inv)) + 1 for k in range(3)]
    if d[0] * d[1] * d[2] > np.iinfo(np.int32).max:
        return p.copy()                                         # "Leaf size is too small": output = input
    min_b = np.floor(mn * inv).astype(np.int32)
    max_b = np.floor(mx * inv).astype(np.int32)
    div_b = max_b - min_b + 1
    mul = np.array([1, div_b[0], div_b[0] * div_b[1]], np.int64)
    ijk = (np.floor(p[:, :3] * inv) - min_b.astype(f32)).astype(np.int32)
    idx = ijk[:, 0].astype(np.int64) * mul[0] + ijk[:, 1] * mul[1] + ijk[:, 2] * mul[2]
    order = np.argsort(idx, kind="stable")
    sidx = idx[order]
    head = np.ones(len(p), bool); head[1:] = sidx[1:] != sidx[:-1]
    vox = np.cumsum(head) - 1
    nv = int(vox[-1]) + 1
    acc = np.zeros((nv, 4), f32)
    np.add.at(acc, vox, p[order])                               # unbuffered: sequential fp32 adds in sorted (= input) order
    cnt = np.bincount(vox, minlength=nv).astype(f32)
    return (acc / cnt[:, None]).astype(f32)


# ------------------------------------------------------------------------------------------------ F16
def calculate_distance(d1, d2):
    """EPSCGeneration::calculateDistance (epscGeneration.cpp:633-660) on two 20 x 80 u8 descriptors.
    Returns (score, best shift i in [-10, 10) or None when no shift beats difference = 1.0)."""
    a = np.asarray(d1, np.uint8).reshape(20, 80).astype(np.int64)
    b = np.asarray(d2, np.uint8).reshape(20, 80).astype(np.int64)
    difference, best = 1.0, None
    for i in range(-10, 10):
        cols = (np.arange(80) + i) % 80
        match_count = int(np.abs(a - b[:, cols]).sum())
        diff_temp = match_count / (80 * 20 * 255)
        if diff_temp < difference:
            difference, best = diff_temp, i
    return 1 - difference, best


# ------------------------------------------------------------------------------------------------
# LOAM feature extraction (round 2): projectPointCloud .. extractFeatures written again from
# src/core/laserProcessing.cpp:467-713, as literally as Python allows (member arrays become numpy arrays, the loops stay
# loops).  Used by tests/test_pyref_oracle.py to pin oracle/orc_features.cpp.  The same resolutions of reference UB as
# listed at the top of orc_features.cpp apply (they are properties of the reference, not of either restatement): entries
# of cloudSmoothness outside the stencil range are {0, i}; std::sort ties are ordered by index; a neighbour walk that would
# read pointColInd outside [0, M) stops.
# ------------------------------------------------------------------------------------------------
def extract_features(pts4, ring, n_scan=64, horizon=1800, downsample_rate=1, min_range=0.0, max_range=70.0, edge_thr=1.0, surf_thr=0.1):
    f32 = np.float32
    p = np.asarray(pts4, f32)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    rng = np.sqrt(x * x + y * y + z * z, dtype=f32)                                   # pointDistance (fp32, left to right)
    row = np.asarray(ring, np.int64)
    ok = ~((rng < f32(min_range)) | (rng > f32(max_range))) & (row >= 0) & (row < n_scan) & (row % downsample_rate == 0)
    # atan2(float, float) is atan2f; taken as the correctly rounded float (see DESIGN.md numerics)
    ang = (np.arctan2(x.astype(np.float64), y.astype(np.float64)).astype(f32) * f32(180)).astype(np.float64) / np.pi
    ang = ang.astype(f32)                                                             # float horizonAngle
    ang_res_x = np.float64(f32(360.0 / float(f32(horizon))))                          # static float ang_res_x = 360.0 / float(Horizon_SCAN)
    t = (ang.astype(np.float64) - 90.0) / ang_res_x
    rnd = np.where(t >= 0, np.floor(t + 0.5), -np.floor(-t + 0.5))                    # round(): halves away from zero
    col = (-rnd + (horizon // 2)).astype(np.int64)                                    # double -> int truncation of an integral value
    col = np.where(col >= horizon, col - horizon, col)
    ok &= (col >= 0) & (col < horizon)
    idx = np.nonzero(ok)[0]
    cell = row[idx] * horizon + col[idx]
    # "if (rangeMat(row, col) != FLT_MAX) continue": the first point to hit a cell keeps it
    ucell, first = np.unique(cell, return_index=True)
    owner = idx[first]                                                                 # ascending cell = row-major = extraction order
    M = len(owner)
    pointColInd = (ucell % horizon).astype(np.int64)
    pointRange = rng[owner]
    rows = ucell // horizon
    start = np.zeros(n_scan, np.int64); end = np.zeros(n_scan, np.int64)
    count = 0
    for i in range(n_scan):
        start[i] = count - 1 + 5
        count += int(np.count_nonzero(rows == i))
        end[i] = count - 1 - 5
    # calculateSmoothness
    curv = np.zeros(M, f32)
    r = pointRange
    for i in range(5, M - 5):
        d = f32(r[i - 5] + r[i - 4]); d = f32(d + r[i - 3]); d = f32(d + r[i - 2]); d = f32(d + r[i - 1]); d = f32(d - f32(r[i] * f32(10)))
        d = f32(d + r[i + 1]); d = f32(d + r[i + 2]); d = f32(d + r[i + 3]); d = f32(d + r[i + 4]); d = f32(d + r[i + 5])
        curv[i] = f32(d * d)
    picked = np.zeros(M, np.int32); label = np.zeros(M, np.int32)
    # markOccludedPoints
    for i in range(5, M - 6):
        depth1, depth2 = r[i], r[i + 1]
        if abs(int(pointColInd[i + 1] - pointColInd[i])) < 10:
            if np.float64(f32(depth1 - depth2)) > 0.3:
                picked[i - 5:i + 1] = 1
            elif np.float64(f32(depth2 - depth1)) > 0.3:
                picked[i + 1:i + 7] = 1
        diff1 = abs(f32(r[i - 1] - r[i])); diff2 = abs(f32(r[i + 1] - r[i]))
        if np.float64(diff1) > 0.02 * np.float64(r[i]) and np.float64(diff2) > 0.02 * np.float64(r[i]):
            picked[i] = 1
    # extractFeatures
    smooth_val = curv.copy(); smooth_ind = np.arange(M)
    corner, sharp, flat, surf = [], [], [], []

    def mark(ind):
        picked[ind] = 1
        for l in range(1, 6):
            if ind + l >= M or ind + l - 1 < 0:
                break
            if abs(int(pointColInd[ind + l] - pointColInd[ind + l - 1])) > 10:
                break
            picked[ind + l] = 1
        for l in range(-1, -6, -1):
            if ind + l < 0 or ind + l + 1 >= M:
                break
            if abs(int(pointColInd[ind + l] - pointColInd[ind + l + 1])) > 10:
                break
            picked[ind + l] = 1

    for i in range(n_scan):
        for j in range(6):
            sp = int((start[i] * (6 - j) + end[i] * j) // 6) if (start[i] * (6 - j) + end[i] * j) >= 0 else -int((-(start[i] * (6 - j) + end[i] * j)) // 6)
            e0 = start[i] * (5 - j) + end[i] * (j + 1)
            ep = (int(e0 // 6) if e0 >= 0 else -int((-e0) // 6)) - 1                 # C integer division truncates toward zero
            if sp >= ep:
                continue
            order = sorted(range(sp, ep), key=lambda k: (smooth_val[k], smooth_ind[k]))   # std::sort over [sp, ep), by_value + index
            vals = [(smooth_val[k], smooth_ind[k]) for k in order]
            for o, k in enumerate(range(sp, ep)):
                smooth_val[k], smooth_ind[k] = vals[o]
            n_pick = 0
            for k in range(ep, sp - 1, -1):
                ind = int(smooth_ind[k])
                if picked[ind] == 0 and curv[ind] > f32(edge_thr):
                    n_pick += 1
                    if n_pick <= 20:
                        label[ind] = 1
                        corner.append(ind)
                        if n_pick <= 4:
                            sharp.append(ind)
                    else:
                        break
                    mark(ind)
            n_pick = 0
            for k in range(sp, ep + 1):
                ind = int(smooth_ind[k])
                if picked[ind] == 0 and curv[ind] < f32(surf_thr):
                    n_pick += 1
                    label[ind] = -1
                    if n_pick <= 10:
                        flat.append(ind)
                    mark(ind)
            for k in range(sp, ep + 1):
                if label[k] <= 0:
                    surf.append(k)
    return {"M": M, "src_index": owner.astype(np.int32), "col_ind": pointColInd.astype(np.int32), "range": pointRange, "curvature": curv,
            "label": label, "start_ring": start.astype(np.int32), "end_ring": end.astype(np.int32),
            "corner_idx": np.array(corner, np.int32), "sharp_idx": np.array(sharp, np.int32), "flat_idx": np.array(flat, np.int32),
            "surf_idx": np.array(surf, np.int32)}


# ------------------------------------------------------------------------------------------------
# EPSC / SEPSC / FEPSC descriptors (epscGeneration.cpp:478-607), numpy restatement pinning oracle/orc_epsc.cpp
# ------------------------------------------------------------------------------------------------
def _epsc_bins(xy):
    f32 = np.float32
    x, y = np.asarray(xy[:, 0], f32), np.asarray(xy[:, 1], f32)
    dist = np.sqrt(x * x + y * y, dtype=f32).astype(np.float64)          # std::sqrt(float) -> float -> double distance
    keep = ~((dist >= 60.0) | (dist < 3.0))
    ring_step = (60.0 - 3.0) / 20
    sector_step = 2 * np.pi / 80
    ring_id = np.floor((dist - 3.0) / ring_step).astype(np.int64)
    angle = np.pi + np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(f32).astype(np.float64)   # atan2f, correctly rounded
    sector_id = np.floor(angle / sector_step).astype(np.int64)
    keep &= (ring_id < 20) & (ring_id >= 0) & (sector_id < 80) & (sector_id >= 0)
    return ring_id, sector_id, keep


def _u8_counts(ring_id, sector_id, sel):
    c = np.zeros((20, 80), np.int64)
    np.add.at(c, (ring_id[sel], sector_id[sel]), 1)
    return c % 256                                                          # unsigned char ++ wraps (Q4)


def epsc_describe(corner4, surf4, sem4, sem_label, lut):
    r, s, k = _epsc_bins(np.asarray(corner4, np.float32)); esc = _u8_counts(r, s, k)
    r, s, k = _epsc_bins(np.asarray(surf4, np.float32)); psc = _u8_counts(r, s, k)
    epsc = ((100 * psc) // (1 + esc)) % 256                                  # int division, narrowed to unsigned char
    r, s, k = _epsc_bins(np.asarray(sem4, np.float32))
    u = np.asarray(lut)[np.asarray(sem_label, np.int64)]
    spsc = _u8_counts(r, s, k & ((u == 40) | (u == 50))); sesc = _u8_counts(r, s, k & (u == 81))
    sepsc = ((100 * spsc) // (1 + sesc)) % 256
    fepsc = np.trunc(sepsc.astype(np.float64) * 0.4 + epsc.astype(np.float64) * 0.6).astype(np.int64) % 256
    return {"epsc": epsc.astype(np.uint8), "sepsc": sepsc.astype(np.uint8), "fepsc": fepsc.astype(np.uint8)}


def loop_project(sem4, sem_label):
    """EPSCGeneration::project (epscGeneration.cpp:84-120): 360 sectors, per sector {count (float ++), x, y, label of the LAST
    point of the sector}; only labels 13, 14, 16, 18, 19; float distance / angle / step as declared upstream."""
    f32 = np.float32
    p = np.asarray(sem4, f32); lab = np.asarray(sem_label, np.int64)
    out = np.zeros((360, 4), f32)
    step = f32(2.0 * np.pi / np.float64(f32(360.0)))                          # float step = 2. * M_PI / sectors_range
    for i in range(len(p)):
        if lab[i] not in (13, 14, 16, 18, 19):
            continue
        x, y = p[i, 0], p[i, 1]
        dist = np.sqrt(f32(x * x + y * y), dtype=f32)
        if np.float64(dist) < 1e-2:
            continue
        angle = f32(np.pi + np.float64(f32(np.arctan2(np.float64(y), np.float64(x)))))   # float angle = M_PI + atan2f(y, x)
        sector = int(np.floor(f32(angle / step)))                              # float / float
        if sector >= 360 or sector < 0:
            continue
        out[sector, 0] = f32(out[sector, 0] + f32(1)); out[sector, 1] = x; out[sector, 2] = y; out[sector, 3] = f32(lab[i])
    return out
