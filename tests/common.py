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
