"""Encode-only profile target (no MPO build): python tools/profile_encode.py [n] [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import qilaplace_b200 as q
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
N = 2**n
dev = torch.device("cuda", 0)
ctx = q.Context(0, stream=torch.cuda.current_stream().cuda_stream)
j = torch.arange(N, dtype=torch.float64, device=dev)
t = j * (1.0 / (2.5 * N))
x = torch.sin(1.0 * t) * torch.exp(-0.08 * t) + torch.sin(2.5 * t) * torch.exp(-0.03 * t)
del j, t
ctx.profile_enable(True)
for it in range(reps):
    ctx.profile_reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    psi = q.signal_mps_dev(ctx, x.data_ptr(), N, False, method="rsvd", **bench.ALGO)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    z = q.ztmps_from_mps(psi, cutoff=bench.ALGO["cutoff"])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("iter", it, "encode ms", (t1 - t0) * 1e3, "split ms", (t2 - t1) * 1e3, "launches", ctx.launch_count())
    for cls, nm in ((0, "stream_gemm"), (3, "qr (outside svd)"), (4, "svd (incl. its qr)")):
        ms, cnt = ctx.profile_read(cls)
        print("   class %-20s %8.3f ms in %d regions" % (nm, ms, cnt))
import numpy as np
idx = (np.arange(64, dtype=np.int64) * 4194301) % N
bits = ((idx[:, None] >> np.arange(n - 1, -1, -1)[None, :]) & 1).astype(np.uint8)
c = q.coefficients(psi, bits)
ref = x[torch.from_numpy(idx).to(dev)].cpu().numpy()
print("ok", psi.bonds, "max |coefficient - x| / max|x| on 64 samples: %.3e" % (np.abs(c - ref).max() / np.abs(ref).max()))
