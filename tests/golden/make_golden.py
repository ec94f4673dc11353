"""Regenerates tests/golden/golden_small.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference ships no golden vectors, so these fixtures
pin the ORACLE's behaviour (and, through the GPU tests, the engine's) against regressions; inputs are re-created
from seeds by lis_slam_b200.synth, only outputs are stored."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lis_slam_b200 import synth  # noqa: E402
from oracle import orc  # noqa: E402


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def inputs():
    sc = synth.Scene(seed=1001)
    m = sc.sample_map(n_edge=6000, n_surf=24000, seed=3001)
    rng = np.random.default_rng(4242)
    truth = synth.random_pose(rng)
    guess = synth.perturb_pose(truth, rng)
    f = sc.sample_scan_features(truth, n_corner=800, n_surf=2400, seed=11)
    sweep = sc.scan(truth, sensor="vlp16", seed=2777)
    return sc, m, truth, guess, f, sweep


def build():
    sc, m, truth, guess, f, sweep = inputs()
    out = {"truth": truth, "guess": guess}
    for v in ("A", "B"):
        pose, res, logs = orc.scan2map(f["corner"], f["surf"], m["corner"], m["surf"], guess, orc.lm_params(v),
                                       clabel=f["corner_label"], slabel=f["surf_label"])
        out["lm_%s_pose" % v] = pose
        out["lm_%s_iters" % v] = np.int32(res.iters)
        out["lm_%s_nsel" % v] = np.array([l.n_sel for l in logs], np.int32)
        out["lm_%s_AtA0" % v] = np.array(logs[0].AtA, np.float32)
        out["lm_%s_poses" % v] = np.array([list(l.pose) for l in logs], np.float32)
    fe = orc.extract_features(sweep["pts"], sweep["ring"], orc.feat_params(n_scan=16))
    out["feat_M"] = np.int32(fe["M"])
    for k in ("corner_idx", "sharp_idx", "flat_idx", "surf_idx", "col_ind", "label"):
        out["feat_%s_sha" % k] = digest(fe[k]); out["feat_%s_n" % k] = np.int32(len(fe[k]))
    out["feat_corner_head"] = fe["corner_idx"][:64]
    ext = sweep["pts"][fe["src_index"]]
    vg = orc.voxel_grid(np.ascontiguousarray(ext[fe["surf_idx"]]), 0.4)
    out["voxel_n"] = np.int32(len(vg)); out["voxel_sha"] = digest(vg); out["voxel_head"] = vg[:16]
    d = orc.epsc_describe(ext[fe["corner_idx"]], ext[fe["surf_idx"]], ext, sweep["label"][fe["src_index"]])
    out["fepsc"] = d["fepsc"]; out["epsc_sha"] = digest(d["epsc"]); out["sepsc_sha"] = digest(d["sepsc"])
    # ---- round 2: the callers either side of the path ----
    # sweep pre-treatment (ring / time synthesis) and constant-velocity de-skew
    pp, pr, pt = orc.pretreat(sweep["pts"], 16, scan_period=0.1, min_range=1.0, max_range=70.0)
    out["pretreat_n"] = np.int32(len(pp)); out["pretreat_pts_sha"] = digest(pp); out["pretreat_ring_sha"] = digest(pr); out["pretreat_time_sha"] = digest(pt)
    dc = orc.deskew_cv(pp[:5000], pt[:5000], 0.1, [8.0, 0.5, 0.0], [0.0, 0.01, 0.3])
    out["deskew_cv_sha"] = digest(dc); out["deskew_cv_head"] = dc[:8]
    # transformUpdate: IMU slerp + clamps
    out["transform_update"] = np.stack([orc.transform_update([0.03, -0.02, 0.5, 1, 2, 3], True, 0.01, 0.04, 0.1, 0.0, 0.0),
                                        orc.transform_update([0.3, -0.2, 0.5, 1, 2, 3], True, 0.1, 1.45, 0.01, 0.25, 2.0),
                                        orc.transform_update([0.3, -0.2, 0.5, 1, 2, 3], False, 0.1, 0.1, 0.01, 0.25, 2.0)])
    # local map: two key frames inserted (class clouds cut from the sweep), sliding-cloud extraction
    cls = [np.ascontiguousarray(ext[c::5]) for c in range(5)]                    # five non-empty class clouds
    sm = orc.Submap()
    c1 = sm.insert(cls, [0, 0, 0.02, 0.5, 0.1, 0.0])
    c2 = sm.insert([c[::2] for c in cls], [0, 0, 0.05, 1.4, 0.2, 0.0], dynrem=(30.0, 0.3, 3.0, 0.03), max_num_pts=2000)
    sc_, ss_, cnt = sm.extract([0, 0, 0.05, 1.4, 0.2, 0.0])
    out["submap_counts"] = np.array(c1 + c2 + cnt, np.int32); out["submap_bound"] = sm.bound.copy()
    out["submap_corner_sha"] = digest(sc_); out["submap_surf_sha"] = digest(ss_); out["submap_n"] = np.array([len(sc_), len(ss_)], np.int32)
    sm.close()
    # ICP verify of a displaced copy against the map (reference parameters)
    src = m["surf"][::3].copy(); src[:, 0] += 0.25; src[:, 1] -= 0.15
    icT, ic = orc.icp(src, m["surf"])
    out["icp_T"] = icT; out["icp_fitness"] = np.float64(ic.fitness); out["icp_iters"] = np.int32(ic.iters)
    # descriptor distance (rotated copy => shift) and the 360-sector projection
    d2 = np.roll(d["fepsc"].reshape(20, 80), 3, axis=1).reshape(-1)
    sc2, sh2 = orc.epsc_distance(d["fepsc"], d2)[:2]
    out["epsc_distance"] = np.array([sc2, sh2], np.float64)
    out["loop_project_sha"] = digest(orc.loop_project(ext, sweep["label"][fe["src_index"]]))
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz"), **build())
    print("written")
