"""Benchmark / test workloads built from the seeded synthetic generator.

`throughput_batch` = BASELINE.json configs[2] shape per GPU: B independent scan-to-map
registrations (HDL-64-shaped feature clouds, ~4k edge + ~12k planar points after voxel
down-sampling) against 200k-point edge/surf local maps (40k + 160k), `n_maps` distinct maps,
`n_scans` distinct scans x B/n_scans distinct initial guesses (truth o perturbation with
|dtheta| <= 0.02 rad, |dt| <= 0.3 m).
"""
import numpy as np

from . import synth


def throughput_batch(B=512, n_maps=8, n_scans=32, n_corner=4000, n_surf=12000, map_edge=40000, map_surf=160000,
                     seed=0, scene=None):
    sc = scene or synth.Scene(seed=1001)
    rng = np.random.default_rng(4001 + 7919 * seed)
    maps = [sc.sample_map(n_edge=map_edge, n_surf=map_surf, seed=3001 + 13 * (seed * n_maps + i)) for i in range(n_maps)]
    scans = []
    for s in range(n_scans):
        truth = synth.random_pose(rng)
        # per-scan sizes vary a little, like real voxel-grid output
        nc = int(n_corner * rng.uniform(0.85, 1.15)); ns = int(n_surf * rng.uniform(0.85, 1.15))
        f = sc.sample_scan_features(truth, n_corner=nc, n_surf=ns, seed=1000 * seed + 100 + s)
        scans.append((f, truth))
    import os
    if os.environ.get("LISREG_BENCH_PRESORT"):      # experiment: queries pre-sorted by map cell (x fastest)
        h = float(os.environ["LISREG_BENCH_PRESORT"])
        for i, (f, truth) in enumerate(scans):
            T = synth.pose_to_T(truth)
            for key in ("corner", "surf"):
                w = f[key][:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
                c = np.floor((w - w.min(0)) / h).astype(np.int64)
                order = np.lexsort((c[:, 0], c[:, 1], c[:, 2]))
                f[key] = np.ascontiguousarray(f[key][order]); f[key + "_label"] = f[key + "_label"][order]
    regs = []
    for b in range(B):
        s = b % n_scans
        f, truth = scans[s]
        guess = synth.perturb_pose(truth, rng)
        regs.append({"scan": s, "map": s % n_maps, "truth": truth, "guess": guess})
    return {"maps": maps, "scans": scans, "regs": regs}


def pack_arena(wl):
    """Packs one private copy of the feature clouds PER REGISTRATION into a contiguous byte arena
    (registrations never share input buffers, so neither the H2D copy nor the L2 is flattered by
    the synthetic generator re-using a scan for several guesses).
    Returns (arena uint8 ndarray, per-registration list of dicts with byte offsets and counts)."""
    chunks, offs, off = [], [], 0
    for r in wl["regs"]:
        f, _ = wl["scans"][r["scan"]]
        o = {}
        for key in ("corner", "surf"):
            a = np.ascontiguousarray(f[key], np.float32)
            o[key] = off; o["n_" + key] = len(a)
            chunks.append(a.view(np.uint8).reshape(-1)); off += a.nbytes
            pad = (-off) % 16
            if pad:
                chunks.append(np.zeros(pad, np.uint8)); off += pad
        offs.append(o)
    return np.concatenate(chunks), offs


def frame_batch(F=256, n_maps=8, n_sweeps=8, map_edge=40000, map_surf=160000, seed=0, scene=None, sensor="hdl64"):
    """BASELINE configs[1]/[2] frames: F raw 64x1800 sweeps (ray-cast HDL-64, range noise 1 cm), each with its
    own local map id and initial guess; `n_sweeps` distinct sweeps x F/n_sweeps distinct guesses."""
    sc = scene or synth.Scene(seed=1001)
    rng = np.random.default_rng(5001 + 7919 * seed)
    maps = [sc.sample_map(n_edge=map_edge, n_surf=map_surf, seed=3001 + 13 * (seed * n_maps + i)) for i in range(n_maps)]
    sweeps = []
    for s in range(n_sweeps):
        truth = synth.random_pose(rng)
        sw = sc.scan(truth, sensor=sensor, seed=2000 + 1000 * seed + s, fast=True)      # C ray-caster: bit-identical to the numpy one
        sweeps.append((sw, truth))
    regs = []
    for b in range(F):
        s = b % n_sweeps
        regs.append({"sweep": s, "map": s % n_maps, "truth": sweeps[s][1], "guess": synth.perturb_pose(sweeps[s][1], rng)})
    return {"maps": maps, "sweeps": sweeps, "regs": regs}


def pack_frame_arena(wl, xyz_only=False):
    """One private copy of the raw sweep (points + ring ids) PER FRAME in a contiguous byte arena.
    xyz_only: 12-byte xyz records instead of packed float4 (lisreg_cloud_layout preset 1: 14 B / point with the ring).
    Returns (arena uint8, [(pts_off, ring_off, n)])."""
    chunks, offs, off = [], [], 0
    for r in wl["regs"]:
        sw, _ = wl["sweeps"][r["sweep"]]
        p = np.ascontiguousarray(sw["pts"][:, :3] if xyz_only else sw["pts"], np.float32); g = np.ascontiguousarray(sw["ring"], np.uint16)
        op = off; chunks.append(p.view(np.uint8).reshape(-1)); off += p.nbytes
        og = off; chunks.append(g.view(np.uint8).reshape(-1)); off += g.nbytes
        pad = (-off) % 16
        if pad:
            chunks.append(np.zeros(pad, np.uint8)); off += pad
        offs.append((op, og, len(p)))
    return np.concatenate(chunks), offs
