import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from lis_slam_b200 import engine as E, synth
from oracle import orc
from common import local_map, reg_case
np.set_printoptions(linewidth=200, precision=6)
eng = E.Engine(0)
m = local_map()
mid = eng.map_create(m["corner"], m["surf"], 1.0)
f, truth, guess = reg_case(0)
po, ro, lo = orc.scan2map(f["corner"], f["surf"], m["corner"], m["surf"], guess, orc.lm_params("A"))
pg, rg, lg = eng.scan2map(mid, f["corner"], f["surf"], guess, E.lm_params("A"), log=True)
print("oracle", ro.status, ro.iters, po)
print("gpu   ", rg.status, rg.iters, pg)
for i in range(max(len(lo), len(lg))):
    if i < len(lo):
        L = lo[i]; print("O", i, L.n_sel, L.n_corner_sel, L.n_surf_sel, L.solved, np.array(L.X), np.array(L.pose))
    if i < len(lg):
        L = lg[i]; print("G", i, L.n_sel, L.n_corner_sel, L.n_surf_sel, L.solved, np.array(L.X), np.array(L.pose))
    if i == 0 and i < len(lo) and i < len(lg):
        print("AtA O\n", np.array(lo[0].AtA).reshape(6, 6)); print("AtA G\n", np.array(lg[0].AtA).reshape(6, 6))
        print("AtB O", np.array(lo[0].AtB)); print("AtB G", np.array(lg[0].AtB))
A = np.array(lg[0].AtA, np.float32).reshape(6, 6); bb = np.array(lg[0].AtB, np.float32)
print("AtB G", bb)
ok, x = orc.qr_solve(A, bb)
print("oracle qr on GPU AtA/AtB:", x)
W, V = orc.jacobi_eigen(A)
print("eig of GPU AtA", W)
print("gpu degenerate", rg.is_degenerate)
