"""Lock-step batch of n=28 signals: time and per-class breakdown. usage: python tools/probe_batch28.py [count]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import qilaplace_b200 as q
count = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = 28
N = 2**n
dev = torch.device("cuda", 0)
ctx = q.Context(0, stream=torch.cuda.current_stream().cuda_stream)
j = torch.arange(N, dtype=torch.float64, device=dev)
t = j * (1.0 / (2.5 * N))
x = torch.sin(1.0 * t) * torch.exp(-0.08 * t) + torch.sin(2.5 * t) * torch.exp(-0.03 * t)
del j, t
xb = x.repeat(count)
ctx.profile_enable(True)
for it in range(3):
    ctx.profile_reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ms = q.signal_mps_batch_dev(ctx, xb.data_ptr(), N, count, False, **bench.ALGO)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    g = ctx.profile_read(0)
    print(f"batch {count} x n={n}: {1e3 * (t1 - t0):.2f} ms ({1e3 * (t1 - t0) / count:.2f} ms/signal), stream gemm {g[0]:.2f} ms in {g[1]} launches; bonds ok {all(m.bonds == ms[0].bonds for m in ms)}")
