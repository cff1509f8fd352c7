"""Key metrics of an `ncu --set full` report, one block per profiled launch (read here, without a GPU).
usage: python tools/ncu_full_summary.py report.ncu-rep [labels...]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
labels = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
KEYS = [("duration (under ncu)", "gpu__time_duration.sum"), ("SM active cycles (avg)", "sm__cycles_active.avg"),
        ("grid", "launch__grid_size"), ("registers/thread", "launch__registers_per_thread"),
        ("tensor pipe active % (of active cycles)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("XU (MUFU) pipe % (of active cycles)", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("DRAM read", "dram__bytes_read.sum"), ("DRAM write", "dram__bytes_write.sum"), ("L2 hit %", "lts__t_sector_hit_rate.pct"),
        ("warp instructions", "smsp__inst_executed.sum")]
stalls = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
for i, r in enumerate(data):
    print(f"== {labels[i] if i < len(labels) else ''}")
    print(f"   kernel: {r[col['Kernel Name']][:120]}")
    for name, k in KEYS:
        if k in col:
            print(f"   {name}: {r[col[k]]} {units[col[k]]}")
    st = []
    for h in stalls:
        try:
            st.append((float(r[col[h]]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
        except ValueError:
            pass
    tot = sum(v for v, _ in st) or 1.0
    print("   stall mix: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:6]))
