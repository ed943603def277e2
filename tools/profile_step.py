"""One warm-up + one step of the bench workload, for ncu captures (never a bench number).
usage: python tools/profile_step.py [n] [coeffs]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import qilaplace_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
B = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
N = 2**n
dev = torch.device("cuda", 0)
ctx = q.Context(0, stream=torch.cuda.current_stream().cuda_stream)
j = torch.arange(N, dtype=torch.float64, device=dev)
t = j * (1.0 / (2.5 * N))
x = torch.sin(1.0 * t) * torch.exp(-0.08 * t) + torch.sin(2.5 * t) * torch.exp(-0.03 * t)
del j, t
bits = torch.from_numpy(bench.hash_bits(B, 2 * n)).to(dev)
out_dev = torch.empty(B, dtype=torch.complex128, device=dev)
scan_a = 10
scan_mode, scan_bits = q.pole_scan_modes(n, 0, 2**n - 2**scan_a, scan_a, scan_a)
scan_dev = torch.empty(4**scan_a, dtype=torch.complex128, device=dev)
W = q.build_zt_mpo(n, bench.OMEGA_R, cutoff=bench.MPO_CUTOFF, maxdim=bench.MPO_MAXDIM, ctx=ctx)
for it in range(2):
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("step%d" % it)
    psi = q.signal_mps_dev(ctx, x.data_ptr(), N, False, method="rsvd", **bench.ALGO)
    z = q.ztmps_from_mps(psi, cutoff=bench.ALGO["cutoff"])
    o = q.apply(W, z)
    q.coefficients_dev(o, bits.data_ptr(), B, out_dev.data_ptr())
    q.coefficient_grid_dev(o, scan_mode, scan_dev.data_ptr(), out_bit=scan_bits)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("ok", psi.bonds, max(o.bonds))
