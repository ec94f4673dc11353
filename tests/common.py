"""Shared seeded scenarios for the parity tests (sizes the CPU oracle finishes in seconds)."""
import functools

import numpy as np

from lis_slam_b200 import synth


@functools.lru_cache(maxsize=None)
def scene():
    return synth.Scene(seed=1001)


@functools.lru_cache(maxsize=None)
def local_map(n_edge=40000, n_surf=160000, seed=3001):
    return scene().sample_map(n_edge=n_edge, n_surf=n_surf, seed=seed)


@functools.lru_cache(maxsize=None)
def reg_case(seed, n_corner=4000, n_surf=12000, rot=0.02, trans=0.3):
    """(features dict, truth pose, initial guess) for one registration."""
    rng = np.random.default_rng(4001 + seed)
    truth = synth.random_pose(rng)
    guess = synth.perturb_pose(truth, rng, rot=rot, trans=trans)
    f = scene().sample_scan_features(truth, n_corner=n_corner, n_surf=n_surf, seed=100 + seed)
    return f, truth, guess


@functools.lru_cache(maxsize=None)
def loop_keyframes(n_out=14, step=2.5, sensor="vlp16"):
    """A there-and-back drive for the loop detector: n_out keyframes along +x every `step` metres, then the way back
    (heading flipped by ~pi, 0.15 m lateral offset, small yaw wobble).  Per keyframe: (corner, surf, semantic cloud,
    labels, odom 4x4) with the clouds in the sensor frame, as EPSCGeneration::loopDetection receives them."""
    from oracle import orc
    rng = np.random.default_rng(77)
    poses = [(step * i, 0.0, 0.02 * rng.standard_normal()) for i in range(n_out)]
    poses += [(step * (n_out - 1 - i) + 0.3 * rng.standard_normal(), 0.15, np.pi + 0.05 * rng.standard_normal()) for i in range(n_out)]
    prm = orc.feat_params(n_scan=16 if sensor == "vlp16" else 64)
    out = []
    for k, (x, y, yaw) in enumerate(poses):
        pose = np.array([0, 0, yaw, x, y, 0], np.float32)
        s = scene().scan(pose, sensor=sensor, seed=6000 + k)
        f = orc.extract_features(s["pts"], s["ring"], prm)
        ext = s["pts"][f["src_index"]]; lab = s["label"][f["src_index"]].astype(np.uint16)
        out.append((np.ascontiguousarray(ext[f["corner_idx"]]), np.ascontiguousarray(ext[f["surf_idx"]]),
                    np.ascontiguousarray(ext), lab, synth.pose_to_T(pose).astype(np.float32)))
    return out
