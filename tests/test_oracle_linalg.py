"""Pins the oracle's small dense routines against the OpenCV wheel (cv2), i.e. against the very
routines the reference calls (cv::eigen odomEstimationNode.cpp:690/:928, cv::solve(QR) :921,
cv::Mat::inv :945).  CPU only."""
import cv2
import numpy as np
import pytest

from oracle import orc


@pytest.mark.parametrize("n", [3, 6])
def test_jacobi_matches_cv_eigen_bit_exact(n):
    rng = np.random.default_rng(n)
    for _ in range(300):
        M = rng.standard_normal((n, n)).astype(np.float32)
        A = (M @ M.T).astype(np.float32) * np.float32(100 if n == 6 else 1)
        _, w, v = cv2.eigen(A)
        W, V = orc.jacobi_eigen(A)
        assert np.array_equal(w.ravel(), W)
        assert np.array_equal(v, V)


def test_qr_solve_matches_cv_solve_bit_exact():
    rng = np.random.default_rng(7)
    for _ in range(300):
        M = rng.standard_normal((6, 6)).astype(np.float32)
        A = (M @ M.T + np.eye(6)).astype(np.float32)
        b = rng.standard_normal(6).astype(np.float32)
        _, x = cv2.solve(A, b.reshape(6, 1), flags=cv2.DECOMP_QR)
        ok, X = orc.qr_solve(A, b)
        assert ok == 1 and np.array_equal(x.ravel(), X)


def test_lu_inverse_matches_cv_invert_bit_exact():
    rng = np.random.default_rng(8)
    for _ in range(300):
        A = rng.standard_normal((6, 6)).astype(np.float32)
        _, inv = cv2.invert(A, flags=cv2.DECOMP_LU)
        ok, I = orc.lu_inv(A)
        assert ok != 0 and np.array_equal(inv, I)


def test_plane_fit_matches_lstsq():
    rng = np.random.default_rng(9)
    for _ in range(300):
        n = rng.standard_normal(3); n /= np.linalg.norm(n)
        d = rng.uniform(1, 30)
        P = rng.uniform(-1, 1, (5, 3)); P -= np.outer(P @ n + d, n); P += rng.normal(0, 0.01, (5, 3))
        P = P.astype(np.float32)
        x = orc.plane_fit(P)
        xr = np.linalg.lstsq(P.astype(np.float64), -np.ones(5), rcond=None)[0]
        assert np.abs(x - xr).max() <= 2e-4 * np.abs(xr).max()


def test_pose_to_affine_is_rz_ry_rx():
    from lis_slam_b200 import synth
    rng = np.random.default_rng(10)
    for _ in range(50):
        p = rng.uniform(-3, 3, 6).astype(np.float32)
        T = orc.pose_to_affine(p)
        Tr = synth.pose_to_T(p)[:3]
        assert np.abs(T - Tr).max() < 1e-6 * max(1, np.abs(Tr).max())


def test_kdtree_matches_brute_force():
    rng = np.random.default_rng(11)
    m = np.zeros((5000, 4), np.float32); m[:, :3] = rng.uniform(-10, 10, (5000, 3))
    q = np.zeros((200, 4), np.float32); q[:, :3] = rng.uniform(-10, 10, (200, 3))
    idx, sqd = orc.knn(m, q, 5)
    d = ((q[:, None, :3] - m[None, :, :3]) ** 2)
    d = (d[..., 0] + d[..., 1]) + d[..., 2]
    ref = np.argsort(d, axis=1, kind="stable")[:, :5]
    assert np.array_equal(idx, ref)
    assert np.array_equal(sqd, np.take_along_axis(d, ref, 1))


def test_corner_and_plane_coefficients_are_geometric():
    """Closed-form check: edge coeff = unit vector from the line to q times s,
    plane coeff = unit normal times s, intensity = s * distance."""
    rng = np.random.default_rng(12)
    nb = np.array([[1, 2, z] for z in (-0.4, -0.2, 0.0, 0.2, 0.4)], np.float32) + rng.normal(0, 1e-3, (5, 3)).astype(np.float32)
    q = np.array([1.3, 2.4, 0.1], np.float32)
    ok, c = orc.corner_coeff(q, nb.ravel())
    dist = 0.5
    s = 1 - 0.9 * dist
    assert ok == 1
    assert np.allclose(c[:3], s * np.array([0.6, 0.8, 0.0]), atol=5e-3)
    assert abs(c[3] - s * dist) < 5e-3
    nb = np.array([[0, 0, -1.73], [0.4, 0, -1.73], [0, 0.4, -1.73], [0.4, 0.4, -1.73], [-0.4, 0.2, -1.73]], np.float32)
    q = np.array([0.1, 0.1, -1.63], np.float32)
    ok, c = orc.surf_coeff(q, nb.ravel())
    assert ok == 1
    r = np.sqrt(np.linalg.norm(q))
    s = 1 - 0.9 * 0.1 / r
    assert np.allclose(np.abs(c[:3]), s * np.array([0, 0, 1.0]), atol=1e-3)
    assert abs(abs(c[3]) - s * 0.1) < 1e-3
