"""CPU tests of the loop-detector restatement (oracle/orc_loop.cpp: project / globalICP / loopDetection,
epscGeneration.cpp:84-120, :258-401, :663-992) against closed-form known answers: the reference ships no golden
vectors for this path, so the oracle is pinned by geometry it must recover."""
import numpy as np

from oracle import orc

from common import loop_keyframes


def test_project_counts_last_point_and_labels():
    # three labelled points in sector 180 (angle pi + atan2(0+, x>0) = pi -> floor(pi / (2pi/360)) = 180), one unlabelled
    pts = np.array([[5, 0.01, 0, 0], [6, 0.02, 1, 0], [7, 0.03, 2, 0], [8, 0.01, 0, 0], [0.001, 0.001, 0, 0]], np.float32)
    lab = np.array([13, 18, 19, 9, 13], np.uint16)          # 9 is not a projected class; the last point is closer than 1 cm
    p = orc.loop_project(pts, lab)
    assert p[:, 0].sum() == 3 and p[180, 0] == 3
    assert p[180, 1] == np.float32(7) and p[180, 2] == np.float32(0.03) and p[180, 3] == 19   # the LAST point of the sector wins
    assert np.count_nonzero(p[:, 3]) == 1


def test_global_icp_recovers_planar_motion():
    rng = np.random.default_rng(5)
    ang = np.sort(rng.uniform(-np.pi, np.pi, 300))
    r = 8 + 4 * np.sin(3 * ang) + rng.uniform(0, 0.05, 300)
    w = np.zeros((300, 4), np.float32); w[:, 0] = r * np.cos(ang); w[:, 1] = r * np.sin(ang)     # landmarks, world frame
    lab = np.full(300, 18, np.uint16)
    yaw, tx, ty = 0.6, 0.4, -0.3                                                                # pose of the current frame in the history frame
    c, s = np.cos(yaw), np.sin(yaw)
    cur = w.copy()
    cur[:, 0] = c * (w[:, 0] - tx) + s * (w[:, 1] - ty); cur[:, 1] = -s * (w[:, 0] - tx) + c * (w[:, 1] - ty)
    T = orc.loop_global_icp(orc.loop_project(w, lab), orc.loop_project(cur, lab), yaw + 0.1)     # odometry yaw guess off by 0.1 rad
    assert abs(np.arctan2(T[1, 0], T[0, 0]) - yaw) < 0.03
    assert abs(T[0, 3] - tx) < 0.25 and abs(T[1, 3] - ty) < 0.25
    assert np.allclose(T[2, :3], [0, 0, 1], atol=1e-6) and np.allclose(T[3], [0, 0, 0, 1])


def test_loop_detection_there_and_back():
    kfs = loop_keyframes()
    n_out = len(kfs) // 2
    det = orc.LoopDetector(use_epsc=True, use_sepsc=True, use_fepsc=True, use_pose=True)
    found = {}
    ncand_total = 0
    for k, (corner, surf, sem, lab, odom) in enumerate(kfs):
        cur, ncand, matches = det.detect(corner, surf, sem, lab, odom)
        assert cur == k
        ncand_total += ncand
        if k < n_out + 4:
            assert ncand == 0 and not matches       # travel gate: > 20 m driven and closer than 1 % of it
        for kind, mid, score, T in matches:
            found.setdefault(kind, []).append((k, mid, score, T))
    det.close()
    assert ncand_total > 0 and "fepsc" in found and "pose" in found
    for kind in ("epsc", "sepsc", "fepsc"):
        for k, mid, score, T in found.get(kind, []):
            assert score > 0.75
            # the matched keyframe is the one passed on the way out: same place, opposite heading
            rel = np.linalg.inv(kfs[mid][4].astype(np.float64)) @ kfs[k][4].astype(np.float64)
            assert np.hypot(rel[0, 3], rel[1, 3]) < 3.0, (kind, k, mid)
            assert np.allclose(T[3], [0, 0, 0, 1]) and T[2, 3] == 0
    # FEPSC (the default descriptor) reports the ICP-refined relative pose: yaw within one descriptor sector, xy within 1 m
    good = 0
    for k, mid, score, T in found["fepsc"]:
        rel = np.linalg.inv(kfs[mid][4].astype(np.float64)) @ kfs[k][4].astype(np.float64)
        dyaw = np.arctan2(T[1, 0], T[0, 0]) - np.arctan2(rel[1, 0], rel[0, 0])
        dyaw = (dyaw + np.pi) % (2 * np.pi) - np.pi
        if abs(dyaw) < np.deg2rad(4.5) and np.hypot(T[0, 3] - rel[0, 3], T[1, 3] - rel[1, 3]) < 1.0:
            good += 1
    assert good >= max(1, len(found["fepsc"]) // 2), (good, len(found["fepsc"]))
