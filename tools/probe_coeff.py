"""Timing probe of the coefficient kernel on the bench's n=28 zT output (never a bench number).
usage: python tools/probe_coeff.py [n] [coeffs] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import qilaplace_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
B = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N = 2**n
dev = torch.device("cuda", 0)
ctx = q.Context(0, stream=torch.cuda.current_stream().cuda_stream)
j = torch.arange(N, dtype=torch.float64, device=dev)
t = j * (1.0 / (2.5 * N))
x = torch.sin(1.0 * t) * torch.exp(-0.08 * t) + torch.sin(2.5 * t) * torch.exp(-0.03 * t)
del j, t
bits = torch.from_numpy(bench.hash_bits(B, 2 * n)).to(dev)
out_dev = torch.empty(B, dtype=torch.complex128, device=dev)
W = q.build_zt_mpo(n, bench.OMEGA_R, cutoff=bench.MPO_CUTOFF, maxdim=bench.MPO_MAXDIM, ctx=ctx)
psi = q.signal_mps_dev(ctx, x.data_ptr(), N, False, method="rsvd", **bench.ALGO)
z = q.ztmps_from_mps(psi, cutoff=bench.ALGO["cutoff"])
o = q.apply(W, z)
bb = [1] + list(o.bonds) + [1]
flops = sum(8.0 * bb[i] * bb[i + 1] for i in range(len(bb) - 1))
q.coefficients_dev(o, bits.data_ptr(), B, out_dev.data_ptr())
torch.cuda.synchronize()
ref = out_dev.clone()
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    q.coefficients_dev(o, bits.data_ptr(), B, out_dev.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("coefficients: %d strings in %.2f ms = %.3f M/s (per 1e6: %.1f ms)%s" % (
        B, ms, B / ms / 1e3, ms * 1e6 / B, ("  %.2f algorithmic TFLOP/s" % (flops * B / ms / 1e9)) if flops else ""))
print("max bond", max(o.bonds), "checksum", complex(out_dev.sum().item()), "repeatable", bool((out_dev == ref).all().item()))
