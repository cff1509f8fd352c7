"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals, shares, averages.
usage: python tools/ncu_summarize.py launches.csv [first_id last_id]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else -1
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
n = 0
for row in csv.DictReader(lines):
    i = int(row["ID"])
    if i < lo or i > hi:
        continue
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
    n += 1
print(f"{n} launches, {tot / 1e3:.3f} ms of kernel time (cold-cache, serialised by ncu)")
print(f"{'us total':>10s} {'share':>7s} {'count':>6s} {'avg us':>8s}  kernel")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} {100 * t / tot:6.1f}% {c:6d} {t / c:8.2f}  {k[:90]}")
