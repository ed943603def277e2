"""Where the zT MPO build goes: python tools/profile_build.py [n]"""
import os, sys, time, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qilaplace_b200 as q
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
ctx = q.Context(0)
W = q.build_zt_mpo(n, 2 * math.pi, cutoff=1e-12, maxdim=128, ctx=ctx)   # warm-up (module load, smem attributes)
for prof in (False, True):
    ctx.profile_reset(); ctx.profile_enable(prof)
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    W = q.build_zt_mpo(n, 2 * math.pi, cutoff=1e-12, maxdim=128, ctx=ctx)
    ctx.sync()
    t1 = time.perf_counter()
    print(f"build_zt_mpo({n}) profile={prof}: {t1 - t0:.3f} s, {ctx.launch_count() - l0} launches, max bond {max(W.bonds)}")
    if prof:
        for cls, nm in ((3, "qr (outside svd)"), (4, "svd (incl. its qr)")):
            ms, cnt = ctx.profile_read(cls)
            print("   class %-20s %9.1f ms in %d regions (%.1f us each)" % (nm, ms, cnt, 1e3 * ms / max(cnt, 1)))
t0 = time.perf_counter(); Wd = q.build_dt_mpo(n, 2 * math.pi, cutoff=1e-12, maxdim=128, ctx=ctx); ctx.sync()
print(f"build_dt_mpo: {time.perf_counter() - t0:.3f} s max bond {max(Wd.bonds)}")
t0 = time.perf_counter(); Wq = q.build_qft_mpo(n, cutoff=1e-12, maxdim=128, ctx=ctx); ctx.sync()
print(f"build_qft_mpo: {time.perf_counter() - t0:.3f} s max bond {max(Wq.bonds)}")
