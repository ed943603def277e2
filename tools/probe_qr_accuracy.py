"""Accuracy of the tall QR on a sketch-like panel (decaying spectrum, numerical rank 5 of 20 columns) against LAPACK."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import qilaplace_b200 as q
rng = np.random.default_rng(7)
for m in (4096, 16384, 640, 2560):
    n = 20
    U0, _ = np.linalg.qr(rng.standard_normal((m, 5)))
    V0, _ = np.linalg.qr(rng.standard_normal((n, 5)))
    s = np.array([1.0, 1e-2, 1e-4, 1e-6, 1e-7])
    Y = (U0 * s) @ V0.T
    Q, R = q.qr(Y, positive=True)
    Qn, Rn = np.linalg.qr(Y)
    sg = np.sign(np.diagonal(Rn)); sg[sg == 0] = 1
    Qn, Rn = Qn * sg, (Rn.T * sg).T
    rec = np.abs(Q @ R - Y).max()
    orth = np.abs(Q.T @ Q - np.eye(n)).max()
    # singular values of Y through R: what the randomized SVD sees
    sv, svn = np.linalg.svd(R, compute_uv=False)[:5], np.linalg.svd(Rn, compute_uv=False)[:5]
    rowerr = [np.abs(R[i] - Rn[i]).max() / np.abs(Rn[i]).max() for i in range(5)]
    proj = [np.abs(Q[:, :k] @ (Q[:, :k].T @ U0) - Qn[:, :k] @ (Qn[:, :k].T @ U0)).max() for k in (1, 3, 5)]
    print("m=%5d rec %.1e orth %.1e | rel err of sigma_1..5 from R: %s | R rows 1..5 vs LAPACK: %s | projector diff k=1,3,5: %s" % (
        m, rec, orth, " ".join("%.1e" % abs(a / b - 1) for a, b in zip(sv, s)), " ".join("%.1e" % e for e in rowerr),
        " ".join("%.1e" % e for e in proj)))
