"""The per-point / per-registration device math of csrc/*.cuh is __host__ __device__; tests/libhostcheck.so is the
same source compiled for the HOST by nvcc (built by __graft_entry__.build()).  CPU-only: it must agree bit-for-bit
with the oracle, which isolates arithmetic parity from everything that needs a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from lis_slam_b200 import synth
from oracle import orc

from common import lattice_map, lattice_queries, local_map, reg_case

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
fp = C.POINTER(C.c_float)


def P(a):
    return a.ctypes.data_as(fp)


@pytest.fixture(scope="module")
def H():
    so = os.path.join(HERE, "libhostcheck.so")
    if not os.path.exists(so):
        import __graft_entry__ as g
        g.build()
    return C.CDLL(so)


def test_small_matrices_bit_exact(H):
    rng = np.random.default_rng(0)
    for t in range(200):
        M = rng.standard_normal((6, 6)).astype(np.float32); A = np.ascontiguousarray((M @ M.T * 100).astype(np.float32))
        b = rng.standard_normal(6).astype(np.float32)
        W = np.zeros(6, np.float32); V = np.zeros((6, 6), np.float32); x = np.zeros(6, np.float32)
        H.hc_jacobi6(P(A), P(W), P(V)); H.hc_qr6(P(A), P(b), P(x))
        Wo, Vo = orc.jacobi_eigen(A); ok, xo = orc.qr_solve(A, b)
        assert np.array_equal(W, Wo) and np.array_equal(V, Vo) and np.array_equal(x, xo)
        A3 = np.ascontiguousarray(A[:3, :3] * np.float32(10.0 ** rng.uniform(-5, 0)))
        if t % 5 == 0:
            A3[0, 1] = A3[1, 0] = 0
        W3 = np.zeros(3, np.float32); V3 = np.zeros((3, 3), np.float32)
        H.hc_jacobi3_reg(P(A3), P(W3), P(V3))          # register-only 3x3 Jacobi used per corner point
        Wo, Vo = orc.jacobi_eigen(A3)
        assert np.array_equal(W3, Wo) and np.array_equal(V3, Vo)


def test_coefficients_bit_exact(H):
    m = local_map(); f, truth, guess = reg_case(0, n_corner=1500, n_surf=4000)
    T = synth.pose_to_T(guess)
    for which, key in ((0, "corner"), (1, "surf")):
        src = f[key]; q = np.zeros((len(src), 4), np.float32)
        q[:, :3] = (src[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        idx, sqd = orc.knn(m[key], q, 5)
        n_acc = 0
        for i in range(0, len(q), 5):
            if sqd[i, 4] >= 1.0:
                continue
            nb = np.ascontiguousarray(m[key][idx[i], :3]).astype(np.float32).ravel(); qq = np.ascontiguousarray(q[i, :3])
            raw = np.zeros(5, np.float32)
            if which == 0:
                ok_o, c_o = orc.corner_coeff(qq, nb); ok_h = H.hc_corner_coeff(P(qq), P(nb), P(raw))
            else:
                ok_o, c_o = orc.surf_coeff(qq, nb); ok_h = H.hc_surf_coeff(P(qq), P(nb), P(raw))
            assert ok_o == ok_h
            if ok_o:
                n_acc += 1
                assert np.array_equal(np.array([raw[4] * raw[k] for k in range(4)], np.float32), c_o)
        assert n_acc > 100


def test_grid_knn_exact_for_any_cell_size(H):
    m = local_map(); f, truth, guess = reg_case(1, n_corner=800, n_surf=2500)
    T = synth.pose_to_T(guess)
    for key in ("corner", "surf"):
        src = f[key]; q = np.zeros((len(src), 4), np.float32)
        q[:, :3] = (src[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
        q[:30, :3] += 500.0; q[30:60, 2] += 3.0
        mp = np.ascontiguousarray(m[key])
        io, so = orc.knn(mp, q, 5)
        for gate, h in ((1.0, 0.6), (1.0, 1.003), (2.0, 0.6), (1.0, 0.31)):
            idx = np.empty((len(q), 5), np.int32); sqd = np.empty((len(q), 5), np.float32)
            H.hc_knn5(mp.ctypes.data_as(C.c_void_p), len(mp), q.ctypes.data_as(C.c_void_p), len(q), C.c_float(h), C.c_float(gate),
                      idx.ctypes.data_as(C.c_void_p), sqd.ctypes.data_as(C.c_void_p))
            inside = so < gate
            assert np.array_equal(np.where(inside, io, -1), idx) and np.array_equal(so[inside], sqd[inside])


def test_grid_knn_tie_order_matches_oracle_on_lattice(H):
    """Maps with exactly equidistant / duplicated points: neighbour INDICES (not only distances) equal the oracle's,
    i.e. bit-equal distances are ordered by original index on both sides (grid.cuh knn_key_less)."""
    m = lattice_map()
    for which, key in ((0, "corner"), (1, "surf")):
        mp = np.ascontiguousarray(m[key]); q = lattice_queries(m, which, n=1500)
        io, so = orc.knn(mp, q, 5)
        assert (so[:, 0] == so[:, 1]).mean() > 0.1 and (so[:, 3] == so[:, 4]).mean() > 0.1     # ties are the rule here
        for gate, h in ((1.0, 0.6), (2.0, 0.6), (1.0, 0.31)):
            idx = np.empty((len(q), 5), np.int32); sqd = np.empty((len(q), 5), np.float32)
            H.hc_knn5(mp.ctypes.data_as(C.c_void_p), len(mp), q.ctypes.data_as(C.c_void_p), len(q), C.c_float(h), C.c_float(gate),
                      idx.ctypes.data_as(C.c_void_p), sqd.ctypes.data_as(C.c_void_p))
            inside = so < gate
            assert np.array_equal(np.where(inside, io, -1), idx) and np.array_equal(so[inside], sqd[inside])
