"""CPU/eager restatement of the reference's layout-conditioned UNet forward.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
(or as the timed CPU baseline), never as the thing that is shipped.

The reference (``/root/reference/GLIGEN``) is pure PyTorch, so the restatement
is plain ``torch.nn.functional`` on a flat ``state_dict`` -- the same F.* op
sequence as the reference modules, so that ``torch.autocast`` takes exactly the
same cast decisions on it (that is what makes it usable as the fp16 oracle on
the GPU box, where ``/root/reference`` does not exist).

Parity pinning: the reference ships no tests/goldens, so
the oracle is pinned against outputs of the reference itself, generated in the
build container by ``tests/gen_golden.py`` and committed under
``tests/golden/`` (see ``tests/test_oracle_golden.py``).

Every function cites the reference file:line it follows (paths relative to
``/root/reference/GLIGEN``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

N_GROUPS = 32


# --------------------------------------------------------------------------- config
def default_unet_config() -> dict:
    """Hyper-parameters of the LayoutLLM-T2I UNet (configs/coco2014.yaml:8-30)."""
    return dict(image_size=64, in_channels=4, out_channels=4, model_channels=320,
                attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
                num_heads=8, transformer_depth=1, context_dim=768, fuser_type="gatedSA",
                grounding_in_dim=768, grounding_out_dim=768, fourier_freqs=8)


def block_plan(cfg: dict):
    """Static layer list of the UNet: mirrors the constructor loops of
    ldm/modules/diffusionmodules/openaimodel.py:299-389.

    Returns (input_blocks, middle, output_blocks); each block is a list of
    (kind, prefix, meta) with kind in {"conv_in", "res", "st", "down", "up"}.
    """
    mc = cfg["model_channels"]
    mult = cfg["channel_mult"]
    nres = cfg["num_res_blocks"]
    attn_res = set(cfg["attention_resolutions"])
    heads = cfg["num_heads"]
    inputs = [[("conv_in", "input_blocks.0.0", dict(cin=cfg["in_channels"], cout=mc))]]
    chans = [mc]
    ch, ds, idx = mc, 1, 1
    for level, m in enumerate(mult):
        for _ in range(nres):
            layers = [("res", f"input_blocks.{idx}.0", dict(cin=ch, cout=m * mc))]
            ch = m * mc
            if ds in attn_res:
                layers.append(("st", f"input_blocks.{idx}.1", dict(c=ch, heads=heads)))
            inputs.append(layers)
            chans.append(ch)
            idx += 1
        if level != len(mult) - 1:
            inputs.append([("down", f"input_blocks.{idx}.0", dict(c=ch))])
            chans.append(ch)
            idx += 1
            ds *= 2
    middle = [("res", "middle_block.0", dict(cin=ch, cout=ch)),
              ("st", "middle_block.1", dict(c=ch, heads=heads)),
              ("res", "middle_block.2", dict(cin=ch, cout=ch))]
    outputs = []
    oidx = 0
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nres + 1):
            ich = chans.pop()
            layers = [("res", f"output_blocks.{oidx}.0", dict(cin=ch + ich, cout=m * mc))]
            ch = m * mc
            j = 1
            if ds in attn_res:
                layers.append(("st", f"output_blocks.{oidx}.{j}", dict(c=ch, heads=heads)))
                j += 1
            if level and i == nres:
                layers.append(("up", f"output_blocks.{oidx}.{j}", dict(c=ch)))
                ds //= 2
            outputs.append(layers)
            oidx += 1
    return inputs, middle, outputs


# --------------------------------------------------------------------------- small pieces
def timestep_embedding(t: Tensor, dim: int, max_period: float = 10000.0) -> Tensor:
    """[cos | sin] sinusoidal embedding, fp32 (ldm/modules/diffusionmodules/util.py:161-181)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    ang = t[:, None].float() * freqs[None]
    emb = torch.cat([ang.cos(), ang.sin()], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def fourier_embed(x: Tensor, num_freqs: int = 8, temperature: float = 100.0) -> Tensor:
    """Per-frequency [sin(f x) | cos(f x)], f_k = T^(k/num) (util.py:12-26)."""
    bands = temperature ** (torch.arange(num_freqs) / num_freqs)
    parts = []
    for f in bands:
        parts += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(parts, dim=-1)


def lin(sd: SD, p: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def layer_norm(sd: SD, p: str, x: Tensor) -> Tensor:
    w = sd[p + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[p + ".bias"], 1e-5)


def group_norm32(sd: SD, p: str, x: Tensor) -> Tensor:
    """GroupNorm32: fp32 compute, cast back, eps 1e-5 (util.py:211-229)."""
    return F.group_norm(x.float(), N_GROUPS, sd[p + ".weight"], sd[p + ".bias"], 1e-5).type(x.dtype)


def position_net(sd: SD, boxes: Tensor, masks: Tensor, emb: Tensor, num_freqs: int = 8) -> Tensor:
    """Grounding tokens [B,30,768] (ldm/modules/diffusionmodules/text_grounding_net.py:26-43)."""
    m = masks.unsqueeze(-1)
    pos = fourier_embed(boxes, num_freqs)
    emb = emb * m + (1 - m) * sd["position_net.null_positive_feature"].view(1, 1, -1)
    pos = pos * m + (1 - m) * sd["position_net.null_position_feature"].view(1, 1, -1)
    h = torch.cat([emb, pos], dim=-1)
    h = F.silu(lin(sd, "position_net.linears.0", h))
    h = F.silu(lin(sd, "position_net.linears.2", h))
    return lin(sd, "position_net.linears.4", h)


def mha(q: Tensor, k: Tensor, v: Tensor, heads: int) -> Tensor:
    """softmax(q k^T d^-0.5) v per head; heads are contiguous channel chunks
    (ldm/modules/attention.py:127-141 and :164-176)."""
    B, N, HC = q.shape
    M = k.shape[1]
    d = HC // heads
    q = q.view(B, N, heads, d).permute(0, 2, 1, 3).reshape(B * heads, N, d)
    k = k.view(B, M, heads, d).permute(0, 2, 1, 3).reshape(B * heads, M, d)
    v = v.view(B, M, heads, d).permute(0, 2, 1, 3).reshape(B * heads, M, d)
    sim = torch.einsum("bid,bjd->bij", q, k) * (d ** -0.5)
    attn = sim.softmax(dim=-1)
    o = torch.einsum("bij,bjd->bid", attn, v)
    return o.view(B, heads, N, d).permute(0, 2, 1, 3).reshape(B, N, HC)


def self_attention(sd: SD, p: str, x: Tensor, heads: int) -> Tensor:
    """attention.py:159-178."""
    o = mha(lin(sd, p + ".to_q", x), lin(sd, p + ".to_k", x), lin(sd, p + ".to_v", x), heads)
    return lin(sd, p + ".to_out.0", o)


def cross_attention(sd: SD, p: str, x: Tensor, kv: Tensor, heads: int) -> Tensor:
    """attention.py:122-143 (mask is always None on this path)."""
    o = mha(lin(sd, p + ".to_q", x), lin(sd, p + ".to_k", kv), lin(sd, p + ".to_v", kv), heads)
    return lin(sd, p + ".to_out.0", o)


def geglu_ff(sd: SD, p: str, x: Tensor) -> Tensor:
    """Linear C->8C, [value|gate] split, value*gelu_erf(gate), Linear 4C->C (attention.py:38-65)."""
    a, g = lin(sd, p + ".net.0.proj", x).chunk(2, dim=-1)
    return lin(sd, p + ".net.2", a * F.gelu(g))


def gated_self_attention(sd: SD, p: str, x: Tensor, objs: Tensor, heads: int, scale: float) -> Tensor:
    """GatedSelfAttentionDense.forward (attention.py:226-234)."""
    n = x.shape[1]
    o = lin(sd, p + ".linear", objs)
    sa = self_attention(sd, p + ".attn", layer_norm(sd, p + ".norm1", torch.cat([x, o], dim=1)), heads)
    x = x + scale * torch.tanh(sd[p + ".alpha_attn"]) * sa[:, :n, :]
    x = x + scale * torch.tanh(sd[p + ".alpha_dense"]) * geglu_ff(sd, p + ".ff", layer_norm(sd, p + ".norm2", x))
    return x


def box_pixel_rects(boxes: Tensor, masks: Tensor, h: int, w: int):
    """Integer rectangles and validity with the reference's truncation / break
    rules (attention.py:321-346): l=int(x0*w), t=int(y0*h), r=int(min(x1*w,w)),
    b=int(min(y1*h,h)); slot i is used iff i < sum(mask) and l!=r and t!=b, and
    the FIRST unused slot ends the scan for that sample."""
    B, mo, _ = boxes.shape
    nvalid = torch.sum(masks, dim=-1).tolist()
    wt = torch.full((B, mo), w).to(boxes)
    ht = torch.full((B, mo), h).to(boxes)
    l = (boxes[:, :, 0] * w).to(torch.int).tolist()
    t = (boxes[:, :, 1] * h).to(torch.int).tolist()
    r = torch.minimum(boxes[:, :, 2] * w, wt).to(torch.int).tolist()
    b = torch.minimum(boxes[:, :, 3] * h, ht).to(torch.int).tolist()
    rects = []
    for k in range(B):
        row = []
        for i in range(mo):
            if i < nvalid[k] and l[k][i] != r[k][i] and t[k][i] != b[k][i]:
                row.append((t[k][i], b[k][i], l[k][i], r[k][i]))
            else:
                break
        rects.append(row)
    return rects


def relation_fusion(sd: SD, p: str, x: Tensor, relations: Tensor, boxes: Tensor, masks: Tensor,
                    h: int, w: int, heads: int) -> Tensor:
    """RelationCrossAttention.forward (attention.py:315-359), restated without the
    30x replicated [B,30,H,W,C] temporaries: out = LN3(x) + (1/30) sum_i M_i f_i.
    Dtypes follow the reference: `hid` keeps the LayerNorm output dtype, the
    per-slot features are stored in x.dtype (attention.py:332,343)."""
    B, _, C = x.shape
    mo = boxes.shape[1]
    hid = layer_norm(sd, p + ".norm3", x).view(B, h, w, C)
    rects = box_pixel_rects(boxes, masks, h, w)
    feats = torch.zeros((B, mo, C)).to(x)
    for k in range(B):
        for i, (t, b, l, r) in enumerate(rects[k]):
            feats[k, i] = hid[k, t:b, l:r, :].reshape(-1, C).mean(dim=0)
    feats = feats + torch.tanh(sd[p + ".alpha_attn"]) * cross_attention(
        sd, p + ".attn", layer_norm(sd, p + ".norm1", feats), relations, heads)
    feats = feats + torch.tanh(sd[p + ".alpha_dense"]) * geglu_ff(sd, p + ".ff", layer_norm(sd, p + ".norm2", feats))
    acc = torch.zeros_like(hid)
    for k in range(B):
        for i, (t, b, l, r) in enumerate(rects[k]):
            acc[k, t:b, l:r, :] += feats[k, i].to(hid.dtype)
    out = hid + acc / mo
    return out.view(B, h * w, C)


def transformer_block(sd: SD, p: str, x: Tensor, context: Tensor, objs: Tensor, relations: Tensor,
                      boxes: Tensor, masks: Tensor, h: int, w: int, heads: int, scale: float,
                      taps: Optional[dict] = None, tap_prefix: str = "") -> Tensor:
    """BasicTransformerBlock._forward (attention.py:394-402)."""
    def rec(name, v):
        if taps is not None:
            taps[tap_prefix + ":" + name] = v
    rec("proj_in", x)
    x = self_attention(sd, p + ".attn1", layer_norm(sd, p + ".norm1", x), heads) + x
    rec("attn1", x)
    x = gated_self_attention(sd, p + ".fuser", x, objs, heads, scale)
    rec("fuser", x)
    x = (relation_fusion(sd, p + ".rela_fuse", x, relations, boxes, masks, h, w, heads) + x) / 2
    rec("rela", x)
    x = cross_attention(sd, p + ".attn2", layer_norm(sd, p + ".norm2", x), context, heads) + x
    rec("attn2", x)
    x = geglu_ff(sd, p + ".ff", layer_norm(sd, p + ".norm3", x)) + x
    rec("ff", x)
    return x


def spatial_transformer(sd: SD, p: str, x: Tensor, context, objs, relations, boxes, masks,
                        heads: int, scale: float, taps: Optional[dict] = None) -> Tensor:
    """SpatialTransformer.forward (attention.py:436-446); GroupNorm eps 1e-6 (:78-79)."""
    B, C, H, W = x.shape
    y = F.group_norm(x, N_GROUPS, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    y = F.conv2d(y, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    y = y.permute(0, 2, 3, 1).reshape(B, H * W, C)
    y = transformer_block(sd, p + ".transformer_blocks.0", y, context, objs, relations, boxes, masks,
                          H, W, heads, scale, taps, p)
    y = y.reshape(B, H, W, C).permute(0, 3, 1, 2)
    y = F.conv2d(y, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return y + x


def res_block(sd: SD, p: str, x: Tensor, emb: Tensor) -> Tensor:
    """ResBlock._forward without up/down (openaimodel.py:211-231)."""
    h = F.silu(group_norm32(sd, p + ".in_layers.0", x))
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = lin(sd, p + ".emb_layers.1", F.silu(emb)).type(h.dtype)
    h = h + e[:, :, None, None]
    h = F.silu(group_norm32(sd, p + ".out_layers.0", h))
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


# --------------------------------------------------------------------------- the forward
def null_grounding(B: int, max_objs: int, in_dim: int, dtype, device) -> dict:
    """GroundingNetInput.get_null_input (grounding_input/text_layout_tokinzer_input.py:47-62)."""
    z = lambda *s: torch.zeros(*s).type(dtype).to(device)
    return dict(boxes=z(B, max_objs, 4), masks=z(B, max_objs), positive_embeddings=z(B, max_objs, in_dim))


def unet_forward(sd: SD, cfg: dict, inp: dict, scale: float = 1.0,
                 first_conv: Optional[SD] = None, max_objs: int = 30, taps: Optional[dict] = None) -> Tensor:
    """UNetModel.forward (openaimodel.py:413-459).

    `scale` is what set_alpha_scale writes on every GatedSelfAttentionDense
    (txt2img.py:46-50); `first_conv` = {"weight","bias"} substitutes
    input_blocks.0.0 as restore_first_conv_from_SD does (openaimodel.py:393-405).
    `taps`, if given, is filled with named intermediate activations (tests).
    """
    x = inp["x"]
    B = x.shape[0]
    if "grounding_input" in inp:
        g = inp["grounding_input"]
    else:
        g = null_grounding(B, max_objs, cfg.get("grounding_in_dim", 768), inp["context"].dtype, x.device)
    objs = position_net(sd, g["boxes"], g["masks"], g["positive_embeddings"], cfg.get("fourier_freqs", 8))
    t_emb = timestep_embedding(inp["timesteps"], cfg["model_channels"])
    emb = lin(sd, "time_embed.2", F.silu(lin(sd, "time_embed.0", t_emb)))
    context, relations = inp["context"], inp["relations"]
    boxes, masks = g["boxes"], g["masks"]
    if taps is not None:
        taps["objs"], taps["emb"] = objs, emb

    def run(layers, h):
        for kind, p, meta in layers:
            if kind == "conv_in":
                wgt = first_conv["weight"] if first_conv is not None else sd[p + ".weight"]
                bias = first_conv["bias"] if first_conv is not None else sd[p + ".bias"]
                h = F.conv2d(h, wgt, bias, padding=1)
            elif kind == "res":
                h = res_block(sd, p, h, emb)
            elif kind == "st":
                h = spatial_transformer(sd, p, h, context, objs, relations, boxes, masks, meta["heads"], scale, taps)
            elif kind == "down":          # Downsample: conv3x3 stride 2 pad 1 (openaimodel.py:103-114)
                h = F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
            elif kind == "up":            # Upsample: nearest x2 then conv3x3 (openaimodel.py:75-85)
                h = F.interpolate(h, scale_factor=2, mode="nearest")
                h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
            if taps is not None:
                taps[p] = h
        return h

    ins, mid, outs = block_plan(cfg)
    h, skips = x, []
    for layers in ins:
        h = run(layers, h)
        skips.append(h)
    h = run(mid, h)
    for layers in outs:
        h = run(layers, torch.cat([h, skips.pop()], dim=1))
    h = F.silu(group_norm32(sd, "out.0", h))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


# --------------------------------------------------------------------------- synthetic weights / inputs
def state_dict_spec(cfg: dict):
    """(key, shape, kind) for every parameter, in the reference's state_dict
    grammar (checked against the reference module's own
    state_dict() in tests/gen_golden.py)."""
    mc, ctx = cfg["model_channels"], cfg["context_dim"]
    spec = []

    def conv(p, cout, cin, k):
        spec.append((p + ".weight", (cout, cin, k, k), "w")); spec.append((p + ".bias", (cout,), "b"))

    def linear(p, cout, cin, bias=True):
        spec.append((p + ".weight", (cout, cin), "w"))
        if bias:
            spec.append((p + ".bias", (cout,), "b"))

    def norm(p, c):
        spec.append((p + ".weight", (c,), "g")); spec.append((p + ".bias", (c,), "b"))

    def attn(p, c, kdim):
        linear(p + ".to_q", c, c, False); linear(p + ".to_k", c, kdim, False); linear(p + ".to_v", c, kdim, False)
        linear(p + ".to_out.0", c, c)

    def ff(p, c):
        linear(p + ".net.0.proj", 8 * c, c); linear(p + ".net.2", c, 4 * c)

    def res(p, cin, cout):
        norm(p + ".in_layers.0", cin); conv(p + ".in_layers.2", cout, cin, 3)
        linear(p + ".emb_layers.1", cout, 4 * mc)
        norm(p + ".out_layers.0", cout); conv(p + ".out_layers.3", cout, cout, 3)
        if cin != cout:
            conv(p + ".skip_connection", cout, cin, 1)

    def st(p, c):
        norm(p + ".norm", c); conv(p + ".proj_in", c, c, 1)
        t = p + ".transformer_blocks.0"
        attn(t + ".attn1", c, c); ff(t + ".ff", c); attn(t + ".attn2", c, ctx)
        for n in ("norm1", "norm2", "norm3"):
            norm(t + "." + n, c)
        f = t + ".fuser"
        linear(f + ".linear", c, ctx); attn(f + ".attn", c, c); ff(f + ".ff", c)
        norm(f + ".norm1", c); norm(f + ".norm2", c)
        spec.append((f + ".alpha_attn", (), "a")); spec.append((f + ".alpha_dense", (), "a"))
        r = t + ".rela_fuse"
        attn(r + ".attn", c, ctx); ff(r + ".ff", c)
        for n in ("norm1", "norm2", "norm3"):
            norm(r + "." + n, c)
        spec.append((r + ".alpha_attn", (), "a")); spec.append((r + ".alpha_dense", (), "a"))
        conv(p + ".proj_out", c, c, 1)

    linear("time_embed.0", 4 * mc, mc); linear("time_embed.2", 4 * mc, 4 * mc)
    ins, mid, outs = block_plan(cfg)
    for layers in ins + [mid] + outs:
        for kind, p, meta in layers:
            if kind == "conv_in":
                conv(p, meta["cout"], meta["cin"], 3)
            elif kind == "res":
                res(p, meta["cin"], meta["cout"])
            elif kind == "st":
                st(p, meta["c"])
            elif kind == "down":
                conv(p + ".op", meta["c"], meta["c"], 3)
            elif kind == "up":
                conv(p + ".conv", meta["c"], meta["c"], 3)
    norm("out.0", mc); conv("out.2", cfg["out_channels"], mc, 3)
    gin, gout = cfg.get("grounding_in_dim", 768), cfg.get("grounding_out_dim", 768)
    pdim = cfg.get("fourier_freqs", 8) * 8
    linear("position_net.linears.0", 512, gin + pdim); linear("position_net.linears.2", 512, 512)
    linear("position_net.linears.4", gout, 512)
    spec.append(("position_net.null_positive_feature", (gin,), "n"))
    spec.append(("position_net.null_position_feature", (pdim,), "n"))
    return spec


def _key_seed(key: str, seed: int) -> int:
    hsh = 1469598103934665603
    for ch in key.encode():
        hsh = ((hsh ^ ch) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (hsh ^ (seed * 0x9E3779B97F4A7C15)) & 0x7FFFFFFFFFFFFFFF


def synthetic_state_dict(cfg: dict, seed: int = 0, gate: Optional[float] = None) -> SD:
    """Deterministic random weights that do not depend on module construction
    order: each tensor is drawn from its own generator seeded by a hash of its
    key.  Weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (PyTorch's default scale),
    norm gains 1+0.1n, biases 0.02n, gates alternate +0.5/-0.4 unless `gate`
    (the reference initialises every alpha to 0, which
    would hide the gated paths)."""
    sd: SD = {}
    flip = 0
    for key, shape, kind in state_dict_spec(cfg):
        g = torch.Generator().manual_seed(_key_seed(key, seed))
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "g":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "b":
            t = 0.02 * torch.randn(shape, generator=g)
        elif kind == "n":
            t = 0.05 * torch.randn(shape, generator=g)
        else:  # gates
            t = torch.tensor(gate if gate is not None else (0.5 if flip % 2 == 0 else -0.4))
            flip += 1
        sd[key] = t.float()
    return sd


def synthetic_inputs(B: int, H: int, W: int, n_boxes: int, max_objs: int = 30, n_rel: int = 3,
                     max_rel: int = 10, ctx_len: int = 77, ctx_dim: int = 768, seed: int = 1234) -> dict:
    """Seeded synthetic inputs of SURVEY.md section 8(d)."""
    def gen(s):
        return torch.Generator().manual_seed(seed + s)
    x = torch.randn(B, 4, H, W, generator=gen(0))
    context = torch.randn(1, ctx_len, ctx_dim, generator=gen(1)).repeat(B, 1, 1)
    uc = torch.randn(1, ctx_len, ctx_dim, generator=gen(2)).repeat(B, 1, 1)
    relations = torch.zeros(B, max_rel, ctx_dim)
    relations[:, :n_rel] = torch.randn(1, n_rel, ctx_dim, generator=gen(3))
    emb = torch.zeros(B, max_objs, ctx_dim)
    emb[:, :n_boxes] = torch.randn(B, n_boxes, ctx_dim, generator=gen(4))
    boxes = torch.zeros(B, max_objs, 4)
    u = torch.rand(B, n_boxes, 4, generator=gen(5))
    x0, y0 = u[..., 0] * 0.5, u[..., 1] * 0.5
    bw, bh = 0.2 + 0.3 * u[..., 2], 0.2 + 0.3 * u[..., 3]
    boxes[:, :n_boxes] = torch.stack([x0, y0, (x0 + bw).clamp(max=1.0), (y0 + bh).clamp(max=1.0)], dim=-1)
    masks = torch.zeros(B, max_objs)
    masks[:, :n_boxes] = 1
    return dict(x=x, context=context, uc=uc, relations=relations,
                grounding=dict(boxes=boxes, masks=masks, positive_embeddings=emb))


def seeded_randn(shape, seed: int) -> Tensor:
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def seeded_module_sd(shapes: Dict[str, tuple], seed: int, gates=(0.5, -0.4)) -> SD:
    """Per-key seeded weights for a single sub-module given its state_dict
    shapes (module goldens): matrices ~ 0.05 n, norm gains 1 + 0.3 n, vectors
    0.3 n, alpha_attn / alpha_dense = gates."""
    sd: SD = {}
    for k, shp in shapes.items():
        g = torch.Generator().manual_seed(_key_seed(k, seed))
        if k.endswith("alpha_attn"):
            t = torch.tensor(gates[0])
        elif k.endswith("alpha_dense"):
            t = torch.tensor(gates[1])
        elif len(shp) >= 2:
            t = 0.05 * torch.randn(shp, generator=g)
        elif "norm" in k and k.endswith("weight"):
            t = 1.0 + 0.3 * torch.randn(shp, generator=g)
        else:
            t = 0.3 * torch.randn(shp, generator=g)
        sd[k] = t.float()
    return sd
