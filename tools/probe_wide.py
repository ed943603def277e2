"""Wide-sketch probe: encode with k=100 (l=105) real at n=28 and the C3-type complex signal with k=50 (l=55)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import qilaplace_b200 as q
dev = torch.device("cuda", 0)
ctx = q.Context(0, stream=torch.cuda.current_stream().cuda_stream)
for (n, cplx, kw) in ((int(sys.argv[1]) if len(sys.argv) > 1 else 28, False, dict(k=100, p=5, q=2, cutoff=1e-12)),
                      (24, True, dict(k=50, p=5, q=2, cutoff=1e-12, maxdim=128)),
                      (24, True, dict(k=100, p=5, q=2, cutoff=1e-12))):
    N = 2**n
    j = torch.arange(N, dtype=torch.float64, device=dev)
    if cplx:
        a = 1.00015 * complex(torch.cos(torch.tensor(0.002)), torch.sin(torch.tensor(0.002)))
        sc = 2.0 ** (20 - n)
        x = torch.exp(j * sc * torch.log(torch.tensor(a, dtype=torch.complex128, device=dev))) * torch.cos(0.0061 * j * sc)
    else:
        t = j / (2.5 * N)
        x = torch.sin(t) * torch.exp(-0.08 * t) + torch.sin(2.5 * t) * torch.exp(-0.03 * t)
    del j
    ctx.profile_enable(True)
    for it in range(3):
        ctx.profile_reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            psi = q.signal_mps_dev(ctx, x.data_ptr(), N, cplx, method="rsvd", **kw)
        except Exception as e:
            print("n", n, "complex", cplx, kw, "->", type(e).__name__, str(e)[:200])
            break
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms, cnt, by, fl = ctx.profile_read_work(0)
        print(f"n={n} complex={cplx} k={kw['k']}: encode {1e3 * (t1 - t0):.2f} ms; stream gemm {ms:.2f} ms in {cnt} launches, "
              f"{fl / (ms / 1e3) / 1e12 if ms else 0:.1f} TFLOP/s ({fl / (ms / 1e3) / 1e12 / 37.1 if ms else 0:.2f} of 37.1), bonds max {max(psi.bonds)}")
    del x
