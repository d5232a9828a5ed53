"""Pin the oracle's restatement of the OpenCV primitives against the real OpenCV (cv2 4.13
wheel in the build container) bit-for-bit, and against the committed golden fixtures (which
were generated from cv2 by tests/golden/make_linalg_golden.py) where cv2 is absent."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "linalg_cv2_4.13.npz")


def test_golden_fixture_eigen_qr_inv_gemm():
    g = np.load(GOLD)
    for n in (3, 6):
        A = g[f"eig{n}_A"]
        for i in range(A.shape[0]):
            W, V = oracle.cv_eigen(A[i])
            assert np.array_equal(W, g[f"eig{n}_W"][i]) and np.array_equal(V, g[f"eig{n}_V"][i])
    for tag, (m, n) in (("53", (5, 3)), ("66", (6, 6)), ("33", (3, 3))):
        A = g[f"qr{tag}_A"]; b = g[f"qr{tag}_b"]
        for i in range(A.shape[0]):
            ok, x = oracle.cv_solve_qr(A[i], b[i])
            assert np.array_equal(x, g[f"qr{tag}_x"][i])
    for n in (3, 6):
        A = g[f"inv{n}_A"]
        for i in range(A.shape[0]):
            ok, D = oracle.cv_inv(A[i])
            assert np.array_equal(D, g[f"inv{n}_D"][i])
    for i in range(g["gemm_A"].shape[0]):
        assert np.array_equal(oracle.cv_gemm(g["gemm_A"][i], g["gemm_B"][i]), g["gemm_D"][i])


def test_live_cv2_bitexact():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(42)
    for n in (3, 6):
        for _ in range(300):
            B = (rng.standard_normal((n + 2, n)) * rng.choice([0.01, 1, 30])).astype(np.float32)
            A = (B.T @ B).astype(np.float32); A = ((A + A.T) / 2).astype(np.float32)
            W, V = oracle.cv_eigen(A)
            _, W2, V2 = cv2.eigen(A)
            assert np.array_equal(W, W2.ravel()) and np.array_equal(V, V2)
    for (m, n) in ((5, 3), (6, 6), (3, 3)):
        for _ in range(300):
            A = (rng.standard_normal((m, n)) * rng.choice([0.1, 1, 50])).astype(np.float32)
            b = (-np.ones((m, 1)) if m == 5 else rng.standard_normal((m, 1))).astype(np.float32)
            _, x = oracle.cv_solve_qr(A, b)
            _, x2 = cv2.solve(A, b, flags=cv2.DECOMP_QR)
            assert np.array_equal(x, x2.ravel())
    for n in (3, 6):
        for _ in range(300):
            A = rng.standard_normal((n, n)).astype(np.float32)
            _, D = oracle.cv_inv(A)
            _, D2 = cv2.invert(A)
            assert np.array_equal(D, D2)
    for (m, k, n) in ((6, 700, 6), (6, 700, 1), (3, 150, 3), (6, 6, 6), (6, 12000, 6)):
        A = rng.standard_normal((m, k)).astype(np.float32); B = rng.standard_normal((k, n)).astype(np.float32)
        assert np.array_equal(oracle.cv_gemm(A, B), cv2.gemm(A, B, 1, None, 0))


def test_eigen_properties():
    rng = np.random.default_rng(3)
    for n in (3, 6):
        B = rng.standard_normal((n + 3, n)).astype(np.float32)
        A = (B.T @ B).astype(np.float32)
        W, V = oracle.cv_eigen(A)
        assert np.all(np.diff(W) <= 0)                       # descending
        assert np.allclose(V @ V.T, np.eye(n), atol=1e-5)    # rows are orthonormal eigenvectors
        assert np.allclose(V.T @ np.diag(W) @ V, A, atol=1e-4 * np.abs(A).max())
