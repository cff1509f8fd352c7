"""UNet-evaluation micro-benchmark for profiling: full-size random-init UNet, one [cond ; uncond] pair per call.
usage: python tools/profile_forward.py [B_img=1] [latent=64] [iters=10] [alpha=both]
Prints launches per evaluation, host enqueue time, device time and the per-class event profile."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "layoutllm_t2i_b200", "dropin"))
import torch  # noqa: E402

import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lat = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
which = sys.argv[4] if len(sys.argv) > 4 else "both"
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
eng = model.engine(30)
h = bench.synthetic_host_inputs(B, lat, lat, 6, 0, pin=False)
d = {k: v.to(dev) for k, v in h.items()}
ctx = torch.cat([d["context"], d["uc"]])
rel = torch.cat([d["relations"], d["relations"]])
g = dict(boxes=d["boxes"], masks=d["masks"], positive_embeddings=d["text_embeddings"])
eng.set_conditioning(ctx, rel, g, lat, lat, B)
x = torch.cat([d["x"], d["x"]])
t = torch.full((2 * B,), 981.0, device=dev)
for alpha in ((1.0, 0.0) if which == "both" else (float(which),)):
    for _ in range(3):
        eng.forward(x, t, alpha)
    torch.cuda.synchronize()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        eng.forward(x, t, alpha)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    n = (eng.launch_count - l0) / iters
    gf = bench.GF_FWD.get(lat, bench.GF_FWD[64])[0 if alpha else 1] * 2 * B
    ms = e0.elapsed_time(e1) / iters
    print(f"alpha={alpha}: {n:.0f} launches/eval, host enqueue {1e3 * t_host / iters:.2f} ms, device {ms:.2f} ms "
          f"-> {gf / ms:.1f} TFLOP/s algorithmic ({gf:.0f} GF)", flush=True)
    if os.environ.get("LTT_CUPROF"):      # ncu --profile-from-start off: capture exactly one evaluation
        torch.cuda.profiler.start()
        eng.forward(x, t, alpha)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        continue
    eng.profile(True)
    eng.forward(x, t, alpha)
    rep = eng.profile_report()
    eng.profile(False)
    for k, v in rep.items():
        tf = f"{v['flops'] / (v['ms'] * 1e-3) / 1e12:.1f} TF/s" if v["flops"] and v["ms"] else ""
        print(f"   {k:10s} {v['ms']:8.3f} ms  {v['launches']:5d} launches  {tf}")
