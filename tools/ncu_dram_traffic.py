"""Aggregate an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` capture of ONE UNet evaluation pair into
per-kernel-class DRAM traffic (bench.py reads profiles/r02_dram_traffic.json for roofline.traffic).
usage: python tools/ncu_dram_traffic.py capture.csv out.json "<how it was captured>" """
import collections
import csv
import json
import sys

CLASSES = (("gemm_tc", "gemm_tc_kernel"), ("attn_tc", "attn_tc_kernel"), ("groupnorm", "gn_"), ("layernorm", "layernorm_kernel"))
path, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
per = collections.defaultdict(lambda: dict(ids=set(), rd=0.0, wr=0.0))
for row in csv.DictReader(lines):
    name = row["Kernel Name"]
    cls = next((c for c, pat in CLASSES if pat in name), "other")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"].lower()
    v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
    d = per[cls]
    d["ids"].add(row["ID"])
    if "read" in row["Metric Name"]:
        d["rd"] += v
    else:
        d["wr"] += v
res = dict(source=note, note="dram__bytes_read.sum + dram__bytes_write.sum per launch, class average over one [cond ; uncond] evaluation pair",
           per_launch_bytes={}, per_eval={})
for cls, d in per.items():
    n = len(d["ids"])
    res["per_launch_bytes"][cls] = round((d["rd"] + d["wr"]) / max(n, 1))
    res["per_eval"][cls] = dict(launches=n, dram_read_MB=round(d["rd"] / 1e6, 1), dram_write_MB=round(d["wr"] / 1e6, 1))
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
