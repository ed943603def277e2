"""ncu target: two C2 steps (256 x n=20 lock-step encode + batched QFT apply); the second is the one to read.
usage: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_c2.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import qilaplace_b200 as q
n, count, kw = bench.C2["n"], bench.C2["count"], bench.C2["kw"]
N = 2**n
dev = torch.device("cuda", 0)
ctx = q.Context(0, stream=torch.cuda.current_stream().cuda_stream)
t = torch.arange(N, dtype=torch.float64, device=dev) / (2.5 * N)
b = torch.arange(count, dtype=torch.float64, device=dev)[:, None]
x = (torch.sin((1 + 0.01 * b) * t) * torch.exp(-0.08 * t) + torch.sin((2.5 + 0.01 * b) * t) * torch.exp(-0.03 * t)).contiguous()
Wq = q.build_qft_mpo(n, ctx=ctx, **bench.C2["qft"])
torch.cuda.synchronize()
for it in range(2):
    l0 = ctx.launch_count()
    ms = q.signal_mps_batch_dev(ctx, x.data_ptr(), N, count, False, **kw)
    out = q.apply_batch(Wq, ms)
    torch.cuda.synchronize()
    print("step", it, "launches", ctx.launch_count() - l0, "max bond", max(max(m.bonds) for m in ms))
