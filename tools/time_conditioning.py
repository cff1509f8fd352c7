"""Per-image cost of ltt_set_conditioning (PositionNet, text / grounding / relation K/V, relation fold) vs the sampler.
usage: python tools/time_conditioning.py [B_img=1] [latent=64]"""
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
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
eng = model.engine(30)
h = bench.synthetic_host_inputs(B, lat, lat, 6, 0, pin=False)
d = {k: v.to(dev) for k, v in h.items()}
ctx = torch.cat([d["context"], d["uc"]])
rel = torch.cat([d["relations"], d["relations"]])
g = dict(boxes=d["boxes"], masks=d["masks"], positive_embeddings=d["text_embeddings"])
for _ in range(3):
    eng.set_conditioning(ctx, rel, g, lat, lat, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(10):
    eng.set_conditioning(ctx, rel, g, lat, lat, B)
e1.record()
t_host = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
print(f"set_conditioning: device {e0.elapsed_time(e1) / 10:.3f} ms, host enqueue {1e3 * t_host:.3f} ms per call")
