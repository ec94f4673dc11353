import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from lis_slam_b200 import engine as E, synth
from oracle import orc
from common import local_map, reg_case
eng = E.Engine(0)
m = local_map()
f, truth, guess = reg_case(0)
T = synth.pose_to_T(guess)
for hint in (2.0, 1.0, 0.25):
    mid = eng.map_create(m["corner"], m["surf"], hint)
    for which, key in ((0, "corner"), (1, "surf")):
        src = f[key]; q = np.zeros((len(src), 4), np.float32)
        q[:, :3] = (src[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        io, so = orc.knn(m[key], q, 5)
        for gate in (1.0, 2.0):
            ig, sg = eng.knn5(mid, which, q, gate)
            inside = so < gate
            print("hint", hint, key, "gate", gate, "idx ok", np.array_equal(np.where(inside, io, -1), ig), "sqd ok", np.array_equal(so[inside], sg[inside]),
                  "nan", np.isnan(sg).sum(), "found5", (ig[:, 4] >= 0).mean(), "oracle", inside[:, 4].mean())
