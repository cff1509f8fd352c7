"""Model-level parity checks: the whole UNet forward / PLMS loop through the C-ABI (layoutllm_t2i_b200.engine)
against the oracle (oracle/*.py) and the committed reference goldens.  Used by tests/test_model_gpu.py and
tools/gpu_model_check.py."""
from __future__ import annotations

import os

import numpy as np
import torch

from layoutllm_t2i_b200.engine import Engine
from oracle import plms_oracle as po
from oracle import unet_oracle as uo

GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"

TINY = dict(image_size=16, in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[2, 1],
            num_res_blocks=1, channel_mult=[1, 2], num_heads=8, transformer_depth=1, context_dim=768,
            fuser_type="gatedSA", grounding_in_dim=768, grounding_out_dim=768, fourier_freqs=8)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def to_dev(d):
    if isinstance(d, dict):
        return {k: to_dev(v) for k, v in d.items()}
    return d.to(DEV) if torch.is_tensor(d) else d


_ENGINES = {}


def engine_for(cfg: dict, seed: int, tag: str = ""):
    key = (repr(sorted((k, repr(v)) for k, v in cfg.items())), seed, tag)
    if key not in _ENGINES:
        sd = uo.synthetic_state_dict(cfg, seed=seed)
        e = Engine(cfg, 0)
        e.load_state_dict(sd)
        e.finalize()
        _ENGINES[key] = (e, sd)
    return _ENGINES[key]


def cfg_batch(syn, B):
    """[cond ; uncond] conditioning of a CFG pair (reference plms.py:116-122)."""
    ctx = torch.cat([syn["context"][:B], syn["uc"][:B]])
    relations = torch.cat([syn["relations"][:B], syn["relations"][:B]])
    return ctx, relations


def oracle_eps(sd, cfg, syn, t, scale, cond, autocast, first_conv=None):
    B = syn["x"].shape[0]
    inp = dict(x=syn["x"], timesteps=torch.full((B,), t, dtype=torch.long, device=syn["x"].device),
               relations=syn["relations"], context=syn["context"] if cond else syn["uc"])
    if cond:
        inp["grounding_input"] = syn["grounding"]
    with torch.no_grad():
        if autocast:
            with torch.autocast("cuda", dtype=torch.float16):
                return uo.unet_forward(sd, cfg, inp, scale=scale, first_conv=first_conv).float()
        from oracle.ref_loader import true_fp32
        with true_fp32():          # no TF32 convolutions in the fp32 oracle
            return uo.unet_forward(sd, cfg, inp, scale=scale, first_conv=first_conv)


def engine_eps_pair(e, syn, t, scale, H, W):
    """cond and uncond eps from ONE [cond ; uncond] engine batch."""
    B = syn["x"].shape[0]
    ctx, relations = cfg_batch(syn, B)
    e.set_conditioning(ctx, relations, syn["grounding"], H, W)
    x2 = torch.cat([syn["x"], syn["x"]])
    tt = torch.full((2 * B,), float(t), device=DEV)
    out = e.forward(x2, tt, scale)
    return out[:B], out[B:]


def check_tiny_vs_reference_golden():
    """Engine (fp16 kernels) against the fp32 outputs of the REFERENCE modules (tests/golden/tiny_unet.pt)."""
    g = torch.load(os.path.join(GOLD, "tiny_unet.pt"), weights_only=False)
    cfg = g["cfg"]
    e, sd = engine_for(cfg, g["seed"])
    syn = uo.synthetic_inputs(**g["syn_args"])
    b, i, box = g["box_override"]
    syn["grounding"]["boxes"][b, i] = torch.tensor(box)
    syn = to_dev(syn)
    worst = 0.0
    for t in (981, 1):
        for s in (1, 0):
            ec, eu = engine_eps_pair(e, syn, t, float(s), 16, 16)
            worst = max(worst, rel(ec, g[f"eps_cond_t{t}_s{s}"]), rel(eu, g[f"eps_unc_t{t}_s{s}"]))
    return worst


def check_unet_vs_oracle(cfg, seed, B, H, W, n_boxes, t, scale, autocast=True, degenerate=False):
    """max over (cond, uncond) of rel-L2(engine, oracle on the GPU [autocast fp16 or fp32])."""
    e, sd = engine_for(cfg, seed)
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    syn = uo.synthetic_inputs(B=B, H=H, W=W, n_boxes=n_boxes, seed=4321)
    if degenerate and n_boxes >= 2:
        syn["grounding"]["boxes"][B - 1, 1] = torch.tensor([0.30, 0.2, 0.3001, 0.9])
    syn = to_dev(syn)
    ec, eu = engine_eps_pair(e, syn, t, scale, H, W)
    oc = oracle_eps(sd_dev, cfg, syn, t, scale, True, autocast)
    ou = oracle_eps(sd_dev, cfg, syn, t, scale, False, autocast)
    return max(rel(ec, oc), rel(eu, ou))


def check_cond_only_batch(cfg, seed, B, H, W, n_boxes, t):
    """A batch with grounding on every row (guidance == 1 path) and one with none."""
    e, sd = engine_for(cfg, seed)
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    syn = to_dev(uo.synthetic_inputs(B=B, H=H, W=W, n_boxes=n_boxes, seed=77))
    tt = torch.full((B,), float(t), device=DEV)
    e.set_conditioning(syn["context"], syn["relations"], syn["grounding"], H, W)
    a = rel(e.forward(syn["x"], tt, 1.0), oracle_eps(sd_dev, cfg, syn, t, 1.0, True, True))
    e.set_conditioning(syn["uc"], syn["relations"], None, H, W)
    b = rel(e.forward(syn["x"], tt, 1.0), oracle_eps(sd_dev, cfg, syn, t, 1.0, False, True))
    return max(a, b)


def check_plms_vs_oracle(cfg, seed, B, H, W, S, guidance=7.5, first_conv_seed=5, per_step_tol=None):
    """Free-running PLMS loop in the library against oracle.plms_sample driving the oracle UNet (autocast fp16).
    Returns rel-L2 of the final latent."""
    e, sd = engine_for(cfg, seed, "plms")     # own engine: the first-conv swap is permanent, as in the reference
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    syn = to_dev(uo.synthetic_inputs(B=B, H=H, W=W, n_boxes=3, seed=555))
    mc = cfg["model_channels"]
    gg = torch.Generator().manual_seed(first_conv_seed)
    fc = dict(weight=(0.2 * torch.randn(mc, 4, 3, 3, generator=gg)).to(DEV), bias=(0.02 * torch.randn(mc, generator=gg)).to(DEV))

    def model_eps(x, t, cond, scale, restored):
        s2 = dict(syn, x=x)
        return oracle_eps(sd_dev, cfg, s2, int(t[0]), float(scale), cond, True, fc if restored else None)

    ref = po.plms_sample(model_eps, syn["x"].clone(), S=S, guidance=guidance)
    ts, a_t, a_prev, s1m = po.plms_tables(S, po.alphas_cumprod())
    ctx, relations = cfg_batch(syn, B)
    if guidance == 1:
        ctx, relations = syn["context"], syn["relations"]
    e.set_conditioning(ctx, relations, syn["grounding"], H, W)
    out = e.plms_sample(syn["x"], ts, a_t, a_prev, s1m, po.alpha_schedule(S), guidance, (fc["weight"], fc["bias"]))
    return rel(out, ref)


# ------------------------------------------------------------------------------------------------ drop-in module tree
def dropin_unet(cfg, seed):
    import sys
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "layoutllm_t2i_b200", "dropin")
    if d not in sys.path:
        sys.path.insert(0, d)
    from grounding_input.text_layout_tokinzer_input import GroundingNetInput
    from ldm.util import instantiate_from_config
    m = instantiate_from_config(dict(
        target="ldm.modules.diffusionmodules.openaimodel.UNetModel",
        params=dict(image_size=cfg["image_size"], in_channels=4, out_channels=4, model_channels=cfg["model_channels"],
                    attention_resolutions=cfg["attention_resolutions"], num_res_blocks=cfg["num_res_blocks"],
                    channel_mult=cfg["channel_mult"], num_heads=8, transformer_depth=1, context_dim=768,
                    fuser_type="gatedSA", use_checkpoint=True,
                    grounding_tokenizer=dict(target="ldm.modules.diffusionmodules.text_grounding_net.PositionNet",
                                             params=dict(in_dim=768, out_dim=768)))))
    sd = uo.synthetic_state_dict(cfg, seed=seed)
    m.load_state_dict(sd, strict=False)           # as reference txt2img.py:106
    m = m.to(DEV).eval()
    m.grounding_tokenizer_input = GroundingNetInput()
    return m, sd


def ref_style_set_alpha_scale(model, alpha_scale):
    """set_alpha_scale as the reference's callers define it (txt2img.py:46-50)."""
    from ldm.modules.attention import GatedCrossAttentionDense, GatedSelfAttentionDense
    for module in model.modules():
        if type(module) == GatedCrossAttentionDense or type(module) == GatedSelfAttentionDense:
            module.scale = alpha_scale


def check_dropin_forward(cfg, seed, B, H, W, t):
    """UNetModel.forward(dict) of the drop-in tree, cond then uncond call as the reference sampler issues them."""
    m, sd = dropin_unet(cfg, seed)
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    syn = to_dev(uo.synthetic_inputs(B=B, H=H, W=W, n_boxes=4, seed=31))
    g = m.grounding_tokenizer_input.prepare(dict(boxes=syn["grounding"]["boxes"], masks=syn["grounding"]["masks"],
                                                 text_embeddings=syn["grounding"]["positive_embeddings"]), None)
    ts = torch.full((B,), t, device=DEV, dtype=torch.long)
    cond = dict(x=syn["x"], timesteps=ts, context=syn["context"], relations=syn["relations"], grounding_input=g,
                inpainting_extra_input=None, grounding_extra_input=None)
    unc = dict(x=syn["x"], timesteps=ts, context=syn["uc"], relations=syn["relations"],
               inpainting_extra_input=None, grounding_extra_input=None)
    worst = 0.0
    for scale in (1.0, 0.0):
        ref_style_set_alpha_scale(m, scale)
        worst = max(worst, rel(m(cond), oracle_eps(sd_dev, cfg, syn, t, scale, True, True)),
                    rel(m(unc), oracle_eps(sd_dev, cfg, syn, t, scale, False, True)))
    return worst


def check_dropin_sampler(cfg, seed, B, H, W, S, mode):
    """PLMSSampler.sample of the drop-in tree (fused or stepwise) against the oracle sampler + oracle UNet."""
    from functools import partial
    from ldm.models.diffusion.ldm import LatentDiffusion
    from ldm.models.diffusion.plms import PLMSSampler
    m, sd = dropin_unet(cfg, seed)
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    syn = to_dev(uo.synthetic_inputs(B=B, H=H, W=W, n_boxes=3, seed=555))
    mc_ = cfg["model_channels"]
    gg = torch.Generator().manual_seed(5)
    fcw, fcb = 0.2 * torch.randn(mc_, 4, 3, 3, generator=gg), 0.02 * torch.randn(mc_, generator=gg)
    m.set_sd_first_conv(fcw, fcb)
    fc = dict(weight=fcw.to(DEV), bias=fcb.to(DEV))

    def model_eps(x, t, cond, scale, restored):
        return oracle_eps(sd_dev, cfg, dict(syn, x=x), int(t[0]), float(scale), cond, True, fc if restored else None)

    ref = po.plms_sample(model_eps, syn["x"].clone(), S=S, guidance=7.5)
    diffusion = LatentDiffusion(linear_start=0.00085, linear_end=0.012, timesteps=1000).to(DEV)
    sampler = PLMSSampler(diffusion, m, alpha_generator_func=partial(po.alpha_schedule, kind=(0.3, 0.0, 0.7)),
                          set_alpha_scale=ref_style_set_alpha_scale)
    sampler.mode = mode
    g = m.grounding_tokenizer_input.prepare(dict(boxes=syn["grounding"]["boxes"], masks=syn["grounding"]["masks"],
                                                 text_embeddings=syn["grounding"]["positive_embeddings"]), None)
    inp = dict(x=syn["x"].clone(), timesteps=None, context=syn["context"], relations=syn["relations"], grounding_input=g,
               inpainting_extra_input=None, grounding_extra_input=None)
    out = sampler.sample(S=S, shape=tuple(syn["x"].shape), input=inp, uc=syn["uc"], guidance_scale=7.5)
    assert m.fuser_scale() == 0.0 and m._sd_conv_active and inp["x"] is out
    return rel(out, ref)


def check_unet_lnfold(cfg, seed, B, H, W, n_boxes, t, scale, mode="1"):
    """The same forward with the LayerNorms folded into the consumer GEMMs (LTT_LNFOLD=1, read at ltt_create); mode "r3": only
    the relation block's norm3 statistics come from the producing GEMM's epilogue."""
    os.environ["LTT_LNFOLD"] = mode
    try:
        e, sd = engine_for(cfg, seed, "lnfold" + mode)
    finally:
        del os.environ["LTT_LNFOLD"]
    sd_dev = {k: v.to(DEV) for k, v in sd.items()}
    syn = to_dev(uo.synthetic_inputs(B=B, H=H, W=W, n_boxes=n_boxes, seed=4321))
    l0 = e.launch_count
    ec, eu = engine_eps_pair(e, syn, t, scale, H, W)
    folded = e.launch_count - l0
    e2, _ = engine_for(cfg, seed)
    l0 = e2.launch_count
    engine_eps_pair(e2, syn, t, scale, H, W)
    assert folded < e2.launch_count - l0, (folded, e2.launch_count - l0)      # the LayerNorm launches are really gone
    return max(rel(ec, oracle_eps(sd_dev, cfg, syn, t, scale, True, True)), rel(eu, oracle_eps(sd_dev, cfg, syn, t, scale, False, True)))


FULL = uo.default_unet_config()

ALL = [
    ("tiny UNet vs reference golden (fp32)", check_tiny_vs_reference_golden, {}, 5e-3),
    ("tiny UNet vs autocast oracle, alpha=1", check_unet_vs_oracle,
     dict(cfg=TINY, seed=7, B=2, H=16, W=16, n_boxes=3, t=981, scale=1.0, degenerate=True), 3e-3),
    ("tiny UNet vs autocast oracle, alpha=0, 24x16 latent", check_unet_vs_oracle,
     dict(cfg=TINY, seed=7, B=3, H=24, W=16, n_boxes=5, t=401, scale=0.0), 3e-3),
    ("tiny UNet with the LayerNorm fold on (LTT_LNFOLD=1), alpha=1", check_unet_lnfold,
     dict(cfg=TINY, seed=7, B=2, H=16, W=16, n_boxes=3, t=981, scale=1.0), 3e-3),
    ("tiny UNet with the LayerNorm fold on, alpha=0", check_unet_lnfold,
     dict(cfg=TINY, seed=7, B=2, H=16, W=16, n_boxes=3, t=401, scale=0.0), 3e-3),
    ("tiny UNet with the norm3 statistics from the producer epilogue (LTT_LNFOLD=r3), alpha=1", check_unet_lnfold,
     dict(cfg=TINY, seed=7, B=2, H=16, W=16, n_boxes=3, t=981, scale=1.0, mode="r3"), 3e-3),
    ("tiny UNet with the norm3 statistics from the producer epilogue, alpha=0", check_unet_lnfold,
     dict(cfg=TINY, seed=7, B=2, H=16, W=16, n_boxes=3, t=401, scale=0.0, mode="r3"), 3e-3),
    ("tiny UNet cond-only / null-only batches", check_cond_only_batch, dict(cfg=TINY, seed=7, B=2, H=16, W=16, n_boxes=30, t=21), 3e-3),
    ("tiny PLMS 5 steps, CFG 7.5", check_plms_vs_oracle, dict(cfg=TINY, seed=7, B=2, H=16, W=16, S=5), 5e-3),
    ("drop-in UNetModel.forward(dict), cond + uncond, both gate values", check_dropin_forward,
     dict(cfg=TINY, seed=7, B=2, H=16, W=16, t=601), 3e-3),
    ("drop-in PLMSSampler.sample fused, 5 steps", check_dropin_sampler, dict(cfg=TINY, seed=7, B=2, H=16, W=16, S=5, mode="fused"), 5e-3),
    ("drop-in PLMSSampler.sample stepwise, 5 steps", check_dropin_sampler, dict(cfg=TINY, seed=7, B=2, H=16, W=16, S=5, mode="stepwise"), 5e-3),
    ("full UNet 64x64 B=1 2 boxes t=981 alpha=1 vs autocast oracle (config 1)", check_unet_vs_oracle,
     dict(cfg=FULL, seed=0, B=1, H=64, W=64, n_boxes=2, t=981, scale=1.0), 3e-3),
    ("full UNet 64x64 B=1 6 boxes t=481 alpha=0 vs autocast oracle", check_unet_vs_oracle,
     dict(cfg=FULL, seed=0, B=1, H=64, W=64, n_boxes=6, t=481, scale=0.0), 3e-3),
    ("full UNet 64x64 B=1 vs fp32 oracle", check_unet_vs_oracle,
     dict(cfg=FULL, seed=0, B=1, H=64, W=64, n_boxes=2, t=981, scale=1.0, autocast=False), 5e-3),
]
