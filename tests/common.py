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


@functools.lru_cache(maxsize=None)
def lattice_map(seed=21):
    """A map made of EXACT ties: a 0.25 m ground lattice, a wall lattice and a pole lattice on binary-exact coordinates,
    every 7th point duplicated (coinciding voxel-grid centroids), the whole array shuffled so that the original index
    order has nothing to do with the spatial order.  Equidistant neighbours are the rule here, so the 5-NN set and
    its order depend on the tie-break: (d^2, original index) on the oracle and on the device."""
    rng = np.random.default_rng(seed)
    g = np.arange(-12.0, 12.0, 0.25, dtype=np.float32)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    ground = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, -1.75, np.float32)], 1)
    wz = np.arange(-1.75, 2.0, 0.25, dtype=np.float32)
    wx, wzz = np.meshgrid(g, wz, indexing="ij")
    wall = np.concatenate([np.stack([wx.ravel(), np.full(wx.size, s * 6.0, np.float32), wzz.ravel()], 1) for s in (-1.0, 1.0)])
    surf = np.concatenate([ground, wall]).astype(np.float32)
    surf = np.concatenate([surf, surf[::7]])                       # duplicated points
    surf = surf[rng.permutation(len(surf))]
    pz = np.arange(-1.75, 2.25, 0.125, dtype=np.float32)
    poles = np.concatenate([np.stack([np.full(len(pz), x, np.float32), np.full(len(pz), y, np.float32), pz], 1)
                            for x in np.arange(-10.0, 10.5, 2.5) for y in (-4.0, 4.0)]).astype(np.float32)
    poles = np.concatenate([poles, poles[::5]])
    poles = poles[rng.permutation(len(poles))]
    mc = np.zeros((len(poles), 4), np.float32); mc[:, :3] = poles
    ms = np.zeros((len(surf), 4), np.float32); ms[:, :3] = surf
    return {"corner": mc, "surf": ms}


def lattice_queries(m, which, n=3000, seed=22):
    """Queries that provoke ties: lattice nodes, cell centres (4 / 8 equidistant nodes), edge midpoints and a few
    random points; all on binary-exact coordinates so the squared distances are bit-equal."""
    rng = np.random.default_rng(seed)
    cloud = m["corner"] if which == 0 else m["surf"]
    base = cloud[rng.choice(len(cloud), n, replace=True), :3].copy()
    step = 0.125 if which == 0 else 0.25
    kind = rng.integers(0, 4, n)
    off = np.zeros((n, 3), np.float32)
    off[kind == 1] = np.float32(step / 2)                                   # cell centre
    off[kind == 2, 0] = np.float32(step / 2)                                # edge midpoint
    off[kind == 3] = rng.uniform(-0.3, 0.3, ((kind == 3).sum(), 3)).astype(np.float32)
    q = np.zeros((n, 4), np.float32); q[:, :3] = base + off
    return q
