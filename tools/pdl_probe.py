"""Does programmatic dependent launch survive graph capture?  Times a chain of small GEMMs eagerly and in a graph, and
dumps the captured graph's edges."""
import math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import ops_checks as oc
a = oc.rn(8192, 320, dtype=torch.float16)
w = oc.rn(320, 320, seed=1, scale=1 / math.sqrt(320)).half()
b = oc.rn(320, seed=2, scale=0.1)
fn = lambda: oc.linear(a, w, b)
for _ in range(5): fn()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): fn()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"eager: host {1e6*(t1-t0)/200:.1f} us/launch, total {1e6*(t2-t0)/200:.1f} us/launch")
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(50): fn()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph: {e0.elapsed_time(e1)/500*1e3:.2f} us/launch")
g.enable_debug_mode() if hasattr(g, "enable_debug_mode") else None
try:
    g2 = torch.cuda.CUDAGraph()
    g2.enable_debug_mode()
    with torch.cuda.graph(g2):
        for _ in range(3): fn()
    g2.debug_dump("/tmp/g.dot")
    txt = open("/tmp/g.dot").read()
    import re
    print("edges:", [l.strip() for l in txt.splitlines() if "->" in l][:6])
    print("programmatic" in txt.lower(), len(txt))
except Exception as e:
    print("dump failed", e)
