"""Which stage produces the sporadic multi-ms gaps: per-step times of 40 back-to-back repetitions of each stage alone."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import qilaplace_b200 as q
n = 28
N = 2**n
dev = torch.device("cuda", 0)
ctx = q.Context(0, stream=torch.cuda.current_stream().cuda_stream)
j = torch.arange(N, dtype=torch.float64, device=dev)
t = j * (1.0 / (2.5 * N))
x = torch.sin(1.0 * t) * torch.exp(-0.08 * t) + torch.sin(2.5 * t) * torch.exp(-0.03 * t)
del j, t
W = q.build_zt_mpo(n, bench.OMEGA_R, cutoff=bench.MPO_CUTOFF, maxdim=bench.MPO_MAXDIM, ctx=ctx)
psi = q.signal_mps_dev(ctx, x.data_ptr(), N, False, method="rsvd", **bench.ALGO)
z = q.ztmps_from_mps(psi, cutoff=bench.ALGO["cutoff"])
keep = {}
def enc(): keep["p"] = q.signal_mps_dev(ctx, x.data_ptr(), N, False, method="rsvd", **bench.ALGO)
def spl(): keep["z"] = q.ztmps_from_mps(psi, cutoff=bench.ALGO["cutoff"])
def app(): keep["o"] = q.apply(W, z)
def loop(tag, fn, cnt=60):
    for _ in range(3): fn()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(cnt + 1)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    evs[0].record()
    host = []
    for i in range(cnt):
        h0 = time.perf_counter()
        fn()
        host.append((time.perf_counter() - h0) * 1e3)
        evs[i + 1].record()
    torch.cuda.synchronize()
    g = [evs[i].elapsed_time(evs[i + 1]) for i in range(cnt)]
    med = sorted(g)[cnt // 2]
    out = [(i, round(g[i], 2), round(host[i], 2)) for i in range(cnt) if g[i] > 1.3 * med + 0.05]
    print(tag, "median %.3f ms; outliers (step, gpu ms, host ms):" % med, out)
loop("encode:", enc)
loop("split :", spl)
loop("apply :", app)
loop("encode:", enc)
def comb():
    h = [time.perf_counter()]
    p = q.signal_mps_dev(ctx, x.data_ptr(), N, False, method="rsvd", **bench.ALGO); h.append(time.perf_counter())
    zz = q.ztmps_from_mps(p, cutoff=bench.ALGO["cutoff"]); h.append(time.perf_counter())
    keep["out"] = q.apply(W, zz); h.append(time.perf_counter())
    del p, zz; h.append(time.perf_counter())
    keep["h"] = [round((h[i + 1] - h[i]) * 1e3, 2) for i in range(4)]
def loop2(cnt=120):
    for _ in range(3): comb()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(cnt + 1)]
    torch.cuda.synchronize()
    evs[0].record()
    hs = []
    for i in range(cnt):
        comb(); hs.append(keep["h"])
        evs[i + 1].record()
    torch.cuda.synchronize()
    g = [evs[i].elapsed_time(evs[i + 1]) for i in range(cnt)]
    med = sorted(g)[cnt // 2]
    print("combined: median %.3f ms; outliers (step, gpu ms, host ms [encode, split, apply, frees]):" % med,
          [(i, round(g[i], 2), hs[i]) for i in range(cnt) if g[i] > 1.05 * med + 0.05], "typical host", hs[50])
loop2()
