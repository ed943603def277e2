"""Timing stability probe: 5 back-to-back steps (encode + split + apply), repeated, with and without an nvidia-smi sampler
running next to it (never a bench number).  usage: python tools/probe_steps.py [reps]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import qilaplace_b200 as q
n = 28
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = 2**n
dev = torch.device("cuda", 0)
ctx = q.Context(0, stream=torch.cuda.current_stream().cuda_stream)
j = torch.arange(N, dtype=torch.float64, device=dev)
t = j * (1.0 / (2.5 * N))
x = torch.sin(1.0 * t) * torch.exp(-0.08 * t) + torch.sin(2.5 * t) * torch.exp(-0.03 * t)
del j, t
W = q.build_zt_mpo(n, bench.OMEGA_R, cutoff=bench.MPO_CUTOFF, maxdim=bench.MPO_MAXDIM, ctx=ctx)
keep = {}
def step():
    psi = q.signal_mps_dev(ctx, x.data_ptr(), N, False, method="rsvd", **bench.ALGO)
    z = q.ztmps_from_mps(psi, cutoff=bench.ALGO["cutoff"])
    keep["out"] = q.apply(W, z)
for _ in range(3):
    step()
torch.cuda.synchronize()
def run(tag):
    res = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 5)
        time.sleep(0.03)
    print(tag, " ".join("%.2f" % v for v in res))
run("no sampler :")
p = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                     stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
time.sleep(0.5)
run("with sampler:")
p.terminate()
run("no sampler :")
# per-step times of 40 back-to-back steps (events recorded per step, one synchronisation at the end)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(41)]
torch.cuda.synchronize()
evs[0].record()
for i in range(40):
    step()
    evs[i + 1].record()
torch.cuda.synchronize()
print("per step   :", " ".join("%.2f" % evs[i].elapsed_time(evs[i + 1]) for i in range(40)))
