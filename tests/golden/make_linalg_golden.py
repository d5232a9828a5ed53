"""Generate tests/golden/linalg_cv2_4.13.npz from the REAL OpenCV (cv2 wheel, 4.13.0, built
without Eigen, Jacobi cv::eigen / Householder cv::solve(DECOMP_QR) / LU cv::invert / cv::gemm).
Run once in the build container:  python tests/golden/make_linalg_golden.py"""
import os

import cv2
import numpy as np

rng = np.random.default_rng(20181001)
out = {}
for n in (3, 6):
    As, Ws, Vs = [], [], []
    for _ in range(64):
        B = (rng.standard_normal((n + 2, n)) * rng.choice([0.01, 1, 30])).astype(np.float32)
        A = (B.T @ B).astype(np.float32); A = ((A + A.T) / 2).astype(np.float32)
        _, W, V = cv2.eigen(A)
        As.append(A); Ws.append(W.ravel()); Vs.append(V)
    out[f"eig{n}_A"] = np.array(As); out[f"eig{n}_W"] = np.array(Ws); out[f"eig{n}_V"] = np.array(Vs)
for tag, (m, n) in (("53", (5, 3)), ("66", (6, 6)), ("33", (3, 3))):
    As, bs, xs = [], [], []
    for t in range(64):
        A = (rng.standard_normal((m, n)) * rng.choice([0.1, 1, 50])).astype(np.float32)
        if m == 5 and t % 2 == 0:     # five near-coplanar map points, as surfOptimization sees them
            nrm = rng.standard_normal(3); nrm /= np.linalg.norm(nrm)
            P = rng.standard_normal((5, 3)) * 0.5
            P -= np.outer(P @ nrm, nrm)
            A = (P + nrm * rng.uniform(1, 40) + rng.standard_normal((5, 3)) * 0.01).astype(np.float32)
        b = (-np.ones((m, 1)) if m == 5 else rng.standard_normal((m, 1))).astype(np.float32)
        _, x = cv2.solve(A, b, flags=cv2.DECOMP_QR)
        As.append(A); bs.append(b.ravel()); xs.append(x.ravel())
    out[f"qr{tag}_A"] = np.array(As); out[f"qr{tag}_b"] = np.array(bs); out[f"qr{tag}_x"] = np.array(xs)
for n in (3, 6):
    As, Ds = [], []
    for _ in range(64):
        A = rng.standard_normal((n, n)).astype(np.float32)
        _, D = cv2.invert(A)
        As.append(A); Ds.append(D)
    out[f"inv{n}_A"] = np.array(As); out[f"inv{n}_D"] = np.array(Ds)
As, Bs, Ds = [], [], []
for _ in range(16):
    A = rng.standard_normal((6, 900)).astype(np.float32); B = rng.standard_normal((900, 6)).astype(np.float32)
    As.append(A); Bs.append(B); Ds.append(cv2.gemm(A, B, 1, None, 0))
out["gemm_A"] = np.array(As); out["gemm_B"] = np.array(Bs); out["gemm_D"] = np.array(Ds)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "linalg_cv2_4.13.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes; cv2", cv2.__version__)
