"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share.
usage: python tools/summarize_launches.py launches.csv [header text ...] > profiles/xxx.txt"""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.OrderedDict()
total = 0.0
n = 0
for row in csv.DictReader(rows):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    us = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
    total += us
    n += 1
for h in sys.argv[2:]:
    print(h)
print(f"launches: {n}, summed kernel time {total / 1000:.3f} ms (cold-cache, serialised under ncu: compare SHARES)")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:72s} n={c:5d} total={t / 1000:10.4f} ms share={t / total:6.4f} avg_us={t / c:10.2f}")
