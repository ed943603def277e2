"""Tensor-core / TMA / barrier mnemonics per kernel of the built library (profiles/*_sass_tensor_tma.txt).
usage: python tools/sass_excerpt.py [round tag] > profiles/rNN_sass_tensor_tma.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "qilaplace.jl_b200", "libqilcuda.so")
tag = sys.argv[1] if len(sys.argv) > 1 else ""
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        c = counts[name]
        c["n"] += 1
        if op.startswith("DMMA"): c["DMMA"] += 1
        if op.startswith("UTMALDG"): c["UTMALDG"] += 1
        if op.startswith("UBLKCP"): c["UBLKCP"] += 1
        if op.startswith("LDGSTS"): c["LDGSTS"] += 1
        if op.startswith("SYNCS"): c["SYNCS"] += 1
        if op.startswith("BAR"): c["BAR"] += 1
print("cuobjdump -sass libqilcuda.so: tensor-core / TMA / barrier mnemonics per kernel (%s); DMMA = DMMA.8x8x4 "
      "(mma.sync m16n8k16 f64 is eight of them), UTMALDG = cp.async.bulk.tensor, UBLKCP = cp.async.bulk" % tag)
for k in sorted(counts):
    c = counts[k]
    if c["DMMA"] or c["UTMALDG"] or c["UBLKCP"] or c["LDGSTS"]:
        print("%-110s instr=%-6d DMMA=%-4d UTMALDG=%d UBLKCP=%d LDGSTS=%d SYNCS(mbarrier)=%d BAR.SYNC=%d" % (
            k, c["n"], c["DMMA"], c["UTMALDG"], c["UBLKCP"], c["LDGSTS"], c["SYNCS"], c["BAR"]))
