#!/usr/bin/env python
"""Small end-to-end workload for `compute-sanitizer` (tests/test_sanitizer_gpu.py): a tiny UNet [cond ; uncond] forward at
both gate values, a 5-step fused PLMS loop (CUDA-graph replay included), a small VAE decode, tiny CLIP text / vision passes,
the image preprocessing and the reward head, all through the C-ABI.
Prints one line per stage; any kernel fault surfaces as a sanitizer error."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import model_checks as mc  # noqa: E402
from layoutllm_t2i_b200.vae import VaeDecoder  # noqa: E402
from oracle import plms_oracle as po  # noqa: E402
from oracle import unet_oracle as uo  # noqa: E402

e, sd = mc.engine_for(mc.TINY, 7)
syn = mc.to_dev(uo.synthetic_inputs(B=2, H=16, W=16, n_boxes=3, seed=4321))
syn["grounding"]["boxes"][1, 1] = torch.tensor([0.30, 0.2, 0.3001, 0.9], device="cuda")
for scale in (1.0, 0.0):
    ec, eu = mc.engine_eps_pair(e, syn, 981, scale, 16, 16)
    torch.cuda.synchronize()
    print(f"forward gate {scale}: finite={bool(torch.isfinite(ec).all() and torch.isfinite(eu).all())}", flush=True)
ts, a_t, a_prev, s1m = po.plms_tables(5, po.alphas_cumprod())
ctx, relations = mc.cfg_batch(syn, 2)
e.set_conditioning(ctx, relations, syn["grounding"], 16, 16)
for _ in range(2):          # second call replays the captured graphs
    out = e.plms_sample(syn["x"], ts, a_t, a_prev, s1m, [1, 1, 0, 0, 0], 7.5, None)
torch.cuda.synchronize()
print(f"plms 5 steps x2: finite={bool(torch.isfinite(out).all())}", flush=True)
g = torch.Generator().manual_seed(0)
dec = VaeDecoder(dict(ch=64, out_ch=3, ch_mult=[1, 2], num_res_blocks=1, z_channels=4, embed_dim=4, scale_factor=0.18215), 0)
vsd = {}
def conv(p, co, ci, k):
    vsd[p + ".weight"] = torch.randn(co, ci, k, k, generator=g) / (ci * k * k) ** 0.5
    vsd[p + ".bias"] = 0.02 * torch.randn(co, generator=g)
def norm(p, c):
    vsd[p + ".weight"] = 1 + 0.1 * torch.randn(c, generator=g)
    vsd[p + ".bias"] = 0.1 * torch.randn(c, generator=g)
def res(p, ci, co):
    norm(p + ".norm1", ci); conv(p + ".conv1", co, ci, 3); norm(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
    if ci != co:
        conv(p + ".nin_shortcut", co, ci, 1)
conv("post_quant_conv", 4, 4, 1); conv("decoder.conv_in", 128, 4, 3)
res("decoder.mid.block_1", 128, 128); res("decoder.mid.block_2", 128, 128)
norm("decoder.mid.attn_1.norm", 128)
for n in ("q", "k", "v", "proj_out"):
    conv("decoder.mid.attn_1." + n, 128, 128, 1)
res("decoder.up.1.block.0", 128, 128); res("decoder.up.1.block.1", 128, 128); conv("decoder.up.1.upsample.conv", 128, 128, 3)
res("decoder.up.0.block.0", 128, 64); res("decoder.up.0.block.1", 64, 64)
norm("decoder.norm_out", 64); conv("decoder.conv_out", 3, 64, 3)
dec.load_state_dict(vsd)
img, u8 = dec.decode(torch.randn(2, 4, 16, 16, generator=g).cuda(), images_u8=True)
torch.cuda.synchronize()
print(f"vae decode: {tuple(img.shape)} finite={bool(torch.isfinite(img).all())} u8={tuple(u8.shape)}", flush=True)
# ---- conditioning prep / reward path: tiny CLIP towers (causal and plain d = 64 attention, quick-GELU epilogue, pooling /
# projection kernels), the image preprocessing kernels on an odd-sized image, the cluster reward head
from layoutllm_t2i_b200.clip import ClipTextEncoder, ClipVisionEncoder  # noqa: E402
from layoutllm_t2i_b200.reward import reward_head  # noqa: E402
from oracle import clip_text_oracle as co  # noqa: E402
from oracle import clip_vision_oracle as cv  # noqa: E402

tcfg, vcfg = co.tiny_clip_text_config(), cv.tiny_clip_vision_config()
text, vision = ClipTextEncoder(tcfg, 0), ClipVisionEncoder(vcfg, 0)
text.load_state_dict(co.random_state_dict(tcfg, seed=1))
vision.load_state_dict(cv.random_state_dict(vcfg, seed=2))
for lengths, L_ in (([5, 0, 12, 75, 1], None), ([3], 5)):
    hid, pooled, emb = text.encode_ids(co.synthetic_ids(tcfg, lengths, L=L_, seed=3), want_embeds=True)
torch.cuda.synchronize()
print(f"clip text: finite={bool(torch.isfinite(hid).all() and torch.isfinite(emb).all())}", flush=True)
imgs = torch.randint(0, 256, (3, 97, 131, 3), dtype=torch.uint8, generator=g).cuda()
px = vision.preprocess(imgs)
_, vp, ve = vision.encode(px)
torch.cuda.synchronize()
print(f"clip vision + preprocess: {tuple(px.shape)} finite={bool(torch.isfinite(px).all() and torch.isfinite(ve).all())}", flush=True)
aes = {k: v.cuda() for k, v in cv.aesthetic_state_dict(vcfg["projection_dim"], seed=4).items()}
r, c, a = reward_head(emb[:3], ve, ve.flip(0), aes, torch.rand(3, generator=g).cuda(), None)
torch.cuda.synchronize()
print(f"reward head: finite={bool(torch.isfinite(r).all())}", flush=True)
print("SANITIZE_SMOKE_DONE", flush=True)
