"""C2 throughput probe (BASELINE configs[1]): 256 signals of n = 20, signal_mps(:rsvd, maxdim=64) + QFT apply.
usage: python tools/bench_c2.py [n] [count]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import qilaplace_b200 as q

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
count = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N = 2**n
dev = torch.device("cuda", 0)
ctx = q.Context(0, stream=torch.cuda.current_stream().cuda_stream)
t = torch.arange(N, dtype=torch.float64, device=dev) / (2.5 * N)
b = torch.arange(count, dtype=torch.float64, device=dev)[:, None]
x = torch.sin((1 + 0.01 * b) * t) * torch.exp(-0.08 * t) + torch.sin((2.5 + 0.01 * b) * t) * torch.exp(-0.03 * t)
x = x.contiguous()
W = q.build_qft_mpo(n, cutoff=1e-14, maxdim=128, ctx=ctx)
kw = dict(k=20, p=10, q=0, cutoff=1e-14, maxdim=64)
torch.cuda.synchronize()
for workers in (16,):
    for rep in range(4):
        t0 = time.perf_counter()
        ms = q.signal_mps_batch_dev(ctx, x.data_ptr(), N, count, False, workers=workers, **kw)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        outs = q.apply_batch(W, ms)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
    print(f"workers {workers:3d}: encode {1e3 * (t1 - t0):8.2f} ms ({count * N / (t1 - t0) / 1e9:7.2f} G samples/s), "
          f"apply {1e3 * (t2 - t1):7.2f} ms, bonds max {max(max(m.bonds) for m in ms)}", flush=True)
