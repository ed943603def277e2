"""Where a C2 step (256 x n=20 encode + QFT apply) spends its time on the HOST: per-phase wall clock of the Python side next
to the device time of the same steps (never a bench number).  usage: python tools/probe_c2_host.py"""
import os, sys, time
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
st = {}
acc = {"encode_call": 0.0, "apply_call": 0.0, "release_prev": 0.0}
def step():
    t0 = time.perf_counter()
    ms = q.signal_mps_batch_dev(ctx, x.data_ptr(), N, count, False, **kw)
    t1 = time.perf_counter()
    out = q.apply_batch(Wq, ms)
    t2 = time.perf_counter()
    st["mps"] = ms; st["out"] = out          # drops the previous step's 512 chains
    t3 = time.perf_counter()
    acc["encode_call"] += t1 - t0; acc["apply_call"] += t2 - t1; acc["release_prev"] += t3 - t2
for _ in range(3):
    step()
torch.cuda.synchronize()
for k in acc: acc[k] = 0.0
K = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); w0 = time.perf_counter()
for _ in range(K):
    step()
e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
print("per step: device %.3f ms, wall %.3f ms | host phases: %s" % (e0.elapsed_time(e1) / K, (w1 - w0) / K * 1e3,
      ", ".join("%s %.3f ms" % (k, v / K * 1e3) for k, v in acc.items())))
# encode alone / apply alone, device time
ctx.profile_enable(False)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
enc, app = 0.0, 0.0
for _ in range(K):
    ev[0].record()
    ms = q.signal_mps_batch_dev(ctx, x.data_ptr(), N, count, False, **kw)
    ev[1].record()
    out = q.apply_batch(Wq, ms)
    ev[2].record()
    torch.cuda.synchronize()
    enc += ev[0].elapsed_time(ev[1]); app += ev[1].elapsed_time(ev[2])
print("device time between events: encode %.3f ms, apply %.3f ms" % (enc / K, app / K))
l0 = ctx.launch_count()
ms = q.signal_mps_batch_dev(ctx, x.data_ptr(), N, count, False, **kw); l1 = ctx.launch_count()
out = q.apply_batch(Wq, ms); l2 = ctx.launch_count()
print("launches: encode", l1 - l0, "apply", l2 - l1)
