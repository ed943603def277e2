import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import qilaplace_b200 as q
rng = np.random.default_rng(0)
A = rng.standard_normal((16384, 20))
for _ in range(2):
    Q, R = q.qr(A, positive=True)
print(np.abs(Q.T @ Q - np.eye(20)).max())
B = rng.standard_normal((20, 20))
for _ in range(2):
    U, S, Vh = q.svd_trunc(rng.standard_normal((120, 120)) @ np.diag(np.logspace(0, -8, 120)), cutoff=1e-12)
