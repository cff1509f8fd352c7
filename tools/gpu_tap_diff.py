"""Per-layer error table: engine vs autocast-fp16 oracle vs fp32 oracle (all on the GPU), next to the noise floor of the
fp16 computation itself: `ac~/ac` = the SAME autocast computation with the input latent perturbed below fp16 resolution
(x * (1 + 5e-5 n): about a tenth of the latent values round to the neighbouring fp16 number at the first conv), and `*/w16` =
distance to the fp32 computation with fp16-ROUNDED weights (weight rounding is common to every fp16 implementation, so
this isolates the activation-rounding noise).
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


def oracle(autocast, sdd=None, xin=None):
    sdd = sd_dev if sdd is None else sdd
    xin = syn["x"] if xin is None else xin
    res = []
    for cond in (True, False):
        tp = {}
        inp = dict(x=xin, timesteps=torch.full((B,), t, dtype=torch.long, device="cuda"),
                   relations=syn["relations"], context=syn["context"] if cond else syn["uc"])
        if cond:
            inp["grounding_input"] = syn["grounding"]
        with torch.no_grad():
            if autocast:
                with torch.autocast("cuda", dtype=torch.float16):
                    o = uo.unet_forward(sdd, cfg, inp, scale=scale, taps=tp)
            else:
                from oracle.ref_loader import true_fp32
                with true_fp32():
                    o = uo.unet_forward(sdd, cfg, inp, scale=scale, taps=tp)
        tp["eps"] = o
        res.append(tp)
    return res


ac, f32 = oracle(True), oracle(False)
# weights (and Linear / Conv biases) rounded to fp16 as autocast does at every call; norm parameters and gates stay fp32
sd16 = {k: (v if ("norm" in k or "alpha" in k or k.endswith("layers.0.weight") or k.endswith("layers.0.bias") or k.startswith("out.0"))
            else v.half().float()) for k, v in sd_dev.items()}
w16 = oracle(False, sd16)
xp = syn["x"] * (1 + 5e-5 * torch.randn(syn["x"].shape, generator=torch.Generator().manual_seed(1)).to("cuda"))
acp = oracle(True, None, xp)


def flat(v):
    """oracle tap -> [rows, C] fp32 in the engine's row order (NHWC)"""
    v = v.float()
    if v.dim() == 4:
        v = v.permute(0, 2, 3, 1)
    return v.reshape(-1, v.shape[-1])


def both(r, name):
    return torch.cat([flat(r[0][name]), flat(r[1][name])])


print(f"{'tap':46s} {'dtype(ac)':10s} {'eng/ac':>10s} {'ac~/ac':>10s} {'eng/f32':>10s} {'ac/f32':>10s} {'eng/w16':>10s} {'ac/w16':>10s} {'w16/f32':>10s}")
for name, v in list(taps.items()) + [("eps", out)]:
    if name not in ac[0]:
        continue
    if name == "eps":
        v = v.float().permute(0, 2, 3, 1).reshape(-1, v.shape[1])
    a, f, w, ap = both(ac, name), both(f32, name), both(w16, name), both(acp, name)
    print(f"{name:46s} {str(ac[0][name].dtype)[6:]:10s} {mc.rel(v, a):10.2e} {mc.rel(ap, a):10.2e} {mc.rel(v, f):10.2e} {mc.rel(a, f):10.2e} "
          f"{mc.rel(v, w):10.2e} {mc.rel(a, w):10.2e} {mc.rel(w, f):10.2e}")
