"""Per-layer error table: engine vs autocast-fp16 oracle vs fp32 oracle (all on the GPU).
usage: python tools/gpu_tap_diff.py [tiny|full] [scale] [t]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import model_checks as mc  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
t = int(sys.argv[3]) if len(sys.argv) > 3 else 981
cfg, seed, B, H, W = (mc.TINY, 7, 2, 16, 16) if which == "tiny" else (mc.FULL, 0, 1, 64, 64)
e, sd = mc.engine_for(cfg, seed)
sd_dev = {k: v.to("cuda") for k, v in sd.items()}
syn = mc.to_dev(uo.synthetic_inputs(B=B, H=H, W=W, n_boxes=3, seed=4321))
ctx, rel = mc.cfg_batch(syn, B)
e.set_conditioning(ctx, rel, syn["grounding"], H, W)
x2 = torch.cat([syn["x"], syn["x"]])
out, taps = e.forward_with_taps(x2, torch.full((2 * B,), float(t), device="cuda"), scale)


def oracle(autocast):
    res = []
    for cond in (True, False):
        tp = {}
        inp = dict(x=syn["x"], timesteps=torch.full((B,), t, dtype=torch.long, device="cuda"),
                   relations=syn["relations"], context=syn["context"] if cond else syn["uc"])
        if cond:
            inp["grounding_input"] = syn["grounding"]
        with torch.no_grad():
            if autocast:
                with torch.autocast("cuda", dtype=torch.float16):
                    o = uo.unet_forward(sd_dev, cfg, inp, scale=scale, taps=tp)
            else:
                o = uo.unet_forward(sd_dev, cfg, inp, scale=scale, taps=tp)
        tp["eps"] = o
        res.append(tp)
    return res


ac, f32 = oracle(True), oracle(False)


def flat(v):
    """oracle tap -> [rows, C] fp32 in the engine's row order (NHWC)"""
    v = v.float()
    if v.dim() == 4:
        v = v.permute(0, 2, 3, 1)
    return v.reshape(-1, v.shape[-1])


print(f"{'tap':46s} {'dtype(ac)':10s} {'eng/ac':>10s} {'eng/f32':>10s} {'ac/f32':>10s}")
for name, v in taps.items():
    if name not in ac[0]:
        continue
    a = torch.cat([flat(ac[0][name]), flat(ac[1][name])])
    f = torch.cat([flat(f32[0][name]), flat(f32[1][name])])
    print(f"{name:46s} {str(ac[0][name].dtype)[6:]:10s} {mc.rel(v, a):10.2e} {mc.rel(v, f):10.2e} {mc.rel(a, f):10.2e}")
a = torch.cat([ac[0]["eps"], ac[1]["eps"]]).float()
f = torch.cat([f32[0]["eps"], f32[1]["eps"]]).float()
print(f"{'eps':46s} {'':10s} {mc.rel(out, a):10.2e} {mc.rel(out, f):10.2e} {mc.rel(a, f):10.2e}")
