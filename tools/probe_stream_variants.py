"""Ring variants of the streaming GEMM (QIL_STREAM_VARIANT, qil_sketch.cu) on the n=28 encode, one process each (the
library reads the variable once).  usage: python tools/probe_stream_variants.py [variants ...]   (never a bench number)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for v in (sys.argv[1:] or ["0", "1", "2", "3", "4"]):
    env = dict(os.environ, QIL_STREAM_VARIANT=v)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_encode.py"), "28", "6"], env=env,
                       capture_output=True, text=True)
    enc = [float(l.split()[4]) for l in r.stdout.splitlines() if l.startswith("iter")]
    sg = [float(l.split()[2]) for l in r.stdout.splitlines() if "class stream_gemm" in l]
    ok = [l for l in r.stdout.splitlines() if l.startswith("ok")]
    print("variant", v, "encode ms best %.3f" % min(enc[1:] or [float("nan")]), "stream gemm ms/encode best %.3f" % min(sg[1:] or [float("nan")]),
          "all", " ".join("%.3f" % t for t in sg), "|", ok[0] if ok else ("FAILED: " + r.stderr[-400:]))
