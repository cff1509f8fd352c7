"""Run every model-level parity check on the GPU and print one line per check (does not stop at failures)."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import model_checks as mc  # noqa: E402

only = sys.argv[1] if len(sys.argv) > 1 else ""
print(torch.cuda.get_device_name(0), flush=True)
bad = 0
for name, fn, kw, tol in mc.ALL:
    if only and only not in name:
        continue
    t0 = time.time()
    try:
        err = fn(**kw)
        torch.cuda.synchronize()
        ok = err < tol
    except Exception as e:  # noqa: BLE001
        err, ok = float("nan"), False
        traceback.print_exc(limit=4)
        if "CUDA" in repr(e) or "cuda" in repr(e):
            print(f"FAIL {name}: CUDA error, aborting")
            sys.exit(2)
    bad += (not ok)
    print(f"{'ok  ' if ok else 'FAIL'} {name}: rel={err:.3e} (tol {tol:g}) [{time.time()-t0:.2f}s]", flush=True)
print("failures:", bad)
sys.exit(1 if bad else 0)
