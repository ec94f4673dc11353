"""16 frames through lisreg_frames_batch (k_feat_segments<true>, k_vox_block<true> need F >= 16 / nseg >= 64? no: nseg = 32 -> the
few-cloud forms; the selection kernel is the target here).  For compute-sanitizer racecheck."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from lis_slam_b200 import engine as E
from common import local_map, scene
sc = scene()
m = local_map()
eng = E.Engine(device=0)
mid = eng.map_create(m["corner"], m["surf"], gate_hint=1.0)
sw = sc.scan(np.array([0, 0, 0.1, 1.0, 0.5, 0], np.float32), seed=3100, fast=True)
F = 16
poses, res = eng.frames_batch([(mid, sw["pts"], sw["ring"])] * F, [np.array([0, 0, 0.1, 1.0, 0.5, 0], np.float32)] * F, E.frame_params("A", early_exit=0, max_iters=2))
print("ok", res[0].n_corner, res[0].n_surf, res[0].iters)
