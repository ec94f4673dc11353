import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lis_slam_b200 import engine as E
from oracle import orc
np.set_printoptions(linewidth=200, precision=6)
eng = E.Engine(0)
rng = np.random.default_rng(0)
bad = 0
for t in range(20):
    M = rng.standard_normal((6, 6)).astype(np.float32); A = (M @ M.T * 100).astype(np.float32); b = rng.standard_normal(6).astype(np.float32)
    r = eng.selftest_smallmat(A, b)
    W, V = orc.jacobi_eigen(A); ok, x = orc.qr_solve(A, b); ok2, inv = orc.lu_inv(A)
    W3, V3 = orc.jacobi_eigen(A[:3, :3])
    e1 = np.array_equal(W, r["E"]); e2 = np.array_equal(V, r["V"]); e3 = np.array_equal(x, r["X"]); e4 = np.array_equal(inv, r["inv"])
    e4 = e4 and np.array_equal(W3, r["W3"]) and np.array_equal(V3, r["V3"])
    if not (e1 and e2 and e3 and e4):
        bad += 1
        if bad < 3:
            print("MISMATCH", t, e1, e2, e3, e4); print(W); print(r["E"]); print(x); print(r["X"])
print("bad", bad, "of 20")
