"""Operator-level parity checks: each CUDA kernel, called through the C-ABI, against the torch oracle on the same
seeded inputs.  Used by tests/test_ops_gpu.py (pytest -m gpu) and tools/gpu_check.py (prints every result)."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from layoutllm_t2i_b200 import _lib as L

DEV = "cuda"


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rn(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def linear(a16, w16, bias=None, act=0, res=None, gate=None, out_dtype=torch.float16):
    M, K = a16.shape
    N = w16.shape[0]
    nout = N // 2 if act == 2 else N
    out = torch.empty(M, nout, device=DEV, dtype=out_dtype)
    L.check(L.lib().ltt_op_linear(L.ptr(a16), M, K, a16.stride(0), L.ptr(w16), N, L.ptr(bias), act, L.ptr(res),
                                  0 if res is None or res.dtype == torch.float16 else 1,
                                  0 if res is None else res.stride(0), 1.0 if gate is None else gate,
                                  0 if gate is None else 1, L.ptr(out), 0 if out_dtype == torch.float16 else 1, nout,
                                  L.stream_ptr()), "ltt_op_linear")
    return out


def check_linear(M, N, K, seed=0, bias=True, act=0, res_dtype=None, gate=None, out_dtype=torch.float16):
    a = rn(M, K, seed=seed, dtype=torch.float16)
    w = rn(N, K, seed=seed + 1, scale=1 / math.sqrt(K))
    b = rn(N, seed=seed + 2, scale=0.1) if bias else None
    nout = N // 2 if act == 2 else N
    res = rn(M, nout, seed=seed + 3, dtype=res_dtype) if res_dtype is not None else None
    if act == 2:
        w16 = torch.empty(N, K, device=DEV, dtype=torch.float16)
        L.check(L.lib().ltt_op_pack_geglu(L.ptr(w.contiguous()), N, K, L.ptr(w16), L.stream_ptr()), "pack_geglu")
    else:
        w16 = w.half()
    y = linear(a, w16, b, act, res, gate, out_dtype)
    ref = F.linear(a.float(), w.half().float(), b)
    if act == 1:
        ref = F.silu(ref)
    elif act == 2:
        v, g = ref.chunk(2, dim=-1)
        ref = v * F.gelu(g)
    if gate is not None:
        ref = ref * gate
    if res is not None:
        ref = ref + res.float()
    return rel(y.float(), ref)


def check_conv3x3(B, H, W, C, N, seed=0, rowvec=False):
    x = rn(B, H, W, C, seed=seed, dtype=torch.float16)
    w = rn(N, C, 3, 3, seed=seed + 1, scale=1 / math.sqrt(9 * C))
    b = rn(N, seed=seed + 2, scale=0.1)
    rv = rn(B, N, seed=seed + 3, dtype=torch.float16) if rowvec else None
    wp = torch.empty(N, 9 * C, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_pack_conv3x3(L.ptr(w.contiguous()), N, C, L.ptr(wp), L.stream_ptr()), "pack_conv")
    out = torch.empty(B, H, W, N, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_conv3x3(L.ptr(x), B, H, W, C, L.ptr(wp), N, L.ptr(b), L.ptr(rv), L.ptr(out),
                                   L.stream_ptr()), "conv3x3")
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), b, padding=1)
    if rowvec:
        ref = ref + rv.float()[:, :, None, None]
    return rel(out.float().permute(0, 3, 1, 2), ref)


def dpad_of(d):
    return {40: 64, 80: 128, 160: 192, 8: 64, 16: 64}[d]


def attention_inputs(B, heads, d, nq, nk, seed=0):
    """q,k in the head-padded layout, v transposed; plus the plain [B,n,C] versions for the oracle."""
    C, dp = heads * d, dpad_of(d)
    q = rn(B, nq, C, seed=seed, dtype=torch.float16)
    k = rn(B, nk, C, seed=seed + 1, dtype=torch.float16)
    v = rn(B, nk, C, seed=seed + 2, dtype=torch.float16)
    rows_k = nk + 3
    pitch = (nk + 63) // 64 * 64
    qp = torch.zeros(B, nq, heads * dp, device=DEV, dtype=torch.float16)
    kp = torch.zeros(B, rows_k, heads * dp, device=DEV, dtype=torch.float16)
    qp.view(B, nq, heads, dp)[..., :d] = q.view(B, nq, heads, d)
    kp.view(B, rows_k, heads, dp)[:, :nk, :, :d] = k.view(B, nk, heads, d)
    kp[:, nk:] = 77.0          # stale rows beyond nk must be ignored
    vt = torch.full((B, C, pitch), 55.0, device=DEV, dtype=torch.float16)
    vt[:, :, :nk] = v.transpose(1, 2)
    return q, k, v, qp, kp, vt, rows_k, pitch


def attention_ref(q, k, v, heads):
    B, nq, C = q.shape
    d = C // heads
    qh = q.float().view(B, nq, heads, d).transpose(1, 2)
    kh = k.float().view(B, -1, heads, d).transpose(1, 2)
    vh = v.float().view(B, -1, heads, d).transpose(1, 2)
    p = (qh @ kh.transpose(-1, -2) * d ** -0.5).softmax(-1)
    return (p @ vh).transpose(1, 2).reshape(B, nq, C)


def check_attention(B, heads, d, nq, nk, seed=0, qscale=1.0):
    q, k, v, qp, kp, vt, rows_k, pitch = attention_inputs(B, heads, d, nq, nk, seed)
    if qscale != 1.0:
        q = q * qscale
        qp = qp * qscale
    C = heads * d
    out = torch.zeros(B, nq, C, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_attention(L.ptr(qp), nq, L.ptr(kp), rows_k, L.ptr(vt), pitch, B, heads, d, dpad_of(d), nq,
                                     nk, d ** -0.5, L.ptr(out), C, L.stream_ptr()), "attention")
    return rel(out.float(), attention_ref(q, k, v, heads))


def check_qkv(B, tokens, C, heads, seed=0):
    d = C // heads
    dp = dpad_of(d)
    a = rn(B, tokens, C, seed=seed, dtype=torch.float16)
    w = rn(3 * C, C, seed=seed + 1, scale=1 / math.sqrt(C))
    rows_k = tokens + 30
    pitch = (rows_k + 63) // 64 * 64
    q = torch.zeros(B, tokens, heads * dp, device=DEV, dtype=torch.float16)
    k = torch.zeros(B, rows_k, heads * dp, device=DEV, dtype=torch.float16)
    vt = torch.zeros(B, C, pitch, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_qkv(L.ptr(a), B, tokens, C, L.ptr(w.half()), heads, dp, L.ptr(q), tokens, L.ptr(k), rows_k,
                               L.ptr(vt), pitch, L.stream_ptr()), "qkv")
    ref = F.linear(a.float(), w.half().float())
    rq, rk, rv = ref.split(C, dim=-1)
    e1 = rel(q.view(B, tokens, heads, dp)[..., :d].reshape(B, tokens, C).float(), rq)
    e2 = rel(k.view(B, rows_k, heads, dp)[:, :tokens, :, :d].reshape(B, tokens, C).float(), rk)
    e3 = rel(vt[:, :, :tokens].transpose(1, 2).float(), rv)
    pad_ok = float(q.view(B, tokens, heads, dp)[..., d:].abs().max()) == 0.0
    return max(e1, e2, e3) if pad_ok else 1.0


def check_groupnorm(B, HW, c0, c1, silu, eps, seed=0):
    x0 = rn(B, HW, c0, seed=seed, dtype=torch.float16) + 0.5
    x1 = rn(B, HW, c1, seed=seed + 1, dtype=torch.float16) * 2 if c1 else None
    C = c0 + c1
    g = 1 + 0.1 * rn(C, seed=seed + 2)
    b = 0.1 * rn(C, seed=seed + 3)
    out = torch.empty(B, HW, C, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_groupnorm(L.ptr(x0), c0, L.ptr(x1), c1, B, HW, L.ptr(g), L.ptr(b), eps, int(silu),
                                     L.ptr(out), L.stream_ptr()), "groupnorm")
    x = torch.cat([x0, x1], dim=-1) if c1 else x0
    ref = F.group_norm(x.float().transpose(1, 2), 32, g, b, eps)
    if silu:
        ref = F.silu(ref.half().float())
    return rel(out.float().transpose(1, 2), ref)


def check_layernorm(M, C, dtype, seed=0):
    x = (rn(M, C, seed=seed) * 3 + 1).to(dtype)
    g = 1 + 0.1 * rn(C, seed=seed + 1)
    b = 0.1 * rn(C, seed=seed + 2)
    o16 = torch.empty(M, C, device=DEV, dtype=torch.float16)
    o32 = torch.empty(M, C, device=DEV, dtype=torch.float32)
    L.check(L.lib().ltt_op_layernorm(L.ptr(x), 0 if dtype == torch.float16 else 1, M, C, L.ptr(g), L.ptr(b), 1e-5,
                                     L.ptr(o16), L.ptr(o32), L.stream_ptr()), "layernorm")
    ref = F.layer_norm(x.float(), (C,), g, b, 1e-5)
    return max(rel(o32, ref), rel(o16.float(), ref) / 4)


def check_small_attention(B, nq, nk, heads, d, seed=0):
    """30 x 10 relation cross-attention core against fp32 torch math on the same fp16 inputs."""
    C = heads * d
    q = rn(B, nq, C, seed=seed, dtype=torch.float16)
    k = rn(B, nk, C, seed=seed + 1, dtype=torch.float16)
    v = rn(B, nk, C, seed=seed + 2, dtype=torch.float16)
    out = torch.empty(B, nq, C, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_small_attention(L.ptr(q), L.ptr(k), L.ptr(v), B, nq, nk, heads, d, d ** -0.5, L.ptr(out),
                                           L.stream_ptr()), "small_attention")
    return rel(out.float(), attention_ref(q, k, v, heads))


def check_rela_attn_fused(G, C, heads, nrel, seed=0):
    """norm1 -> to_q -> attention over the relation tokens -> to_out -> gated residual -> norm2 with the projections
    folded into the relation K / V, against the unfolded fp32 torch math on the same fp16 operands."""
    d, mo = C // heads, 30
    feats = (rn(G, mo, C, seed=seed) * 1.5).half()
    feats[:, 7:] = 0                       # padded box slots pool to zero rows
    wq = rn(C, C, seed=seed + 1, scale=1 / math.sqrt(C)).half()
    wo = rn(C, C, seed=seed + 2, scale=1 / math.sqrt(C)).half()
    bo = 0.1 * rn(C, seed=seed + 3)
    kv = rn(G, nrel, 2 * C, seed=seed + 4, dtype=torch.float16)
    g1, b1 = 1 + 0.1 * rn(C, seed=seed + 5), 0.1 * rn(C, seed=seed + 6)
    g2, b2 = 1 + 0.1 * rn(C, seed=seed + 7), 0.1 * rn(C, seed=seed + 8)
    gate = 0.46
    A = torch.empty(G, heads * nrel, C, device=DEV, dtype=torch.float16)
    Bm = torch.empty_like(A)
    L.check(L.lib().ltt_op_rela_fold(L.ptr(wq), L.ptr(wo), L.ptr(kv), G, nrel, heads, d, d ** -0.5, L.ptr(A), L.ptr(Bm),
                                     L.stream_ptr()), "rela_fold")
    f2 = torch.empty(G, mo, C, device=DEV, dtype=torch.float16)
    l2 = torch.empty_like(f2)
    L.check(L.lib().ltt_op_rela_attn_fused(L.ptr(feats), G, mo, C, heads, nrel, L.ptr(A), L.ptr(Bm), L.ptr(bo), gate, L.ptr(g1),
                                           L.ptr(b1), L.ptr(g2), L.ptr(b2), 1e-5, L.ptr(f2), L.ptr(l2), L.stream_ptr()), "rela_attn_fused")
    x = feats.float()
    ln = F.layer_norm(x, (C,), g1, b1, 1e-5)
    q = F.linear(ln, wq.float())
    k, v = kv.float().split(C, dim=-1)
    att = attention_ref(q, k, v, heads)
    ref2 = x + gate * F.linear(att, wo.float(), bo)
    ref_l2 = F.layer_norm(ref2, (C,), g2, b2, 1e-5)
    return max(rel(f2.float(), ref2), rel(l2.float(), ref_l2))


def check_rela_scatter_ln(B, nb, h, w, C, seed=0):
    """scatter of pooled box features + (out + x) / 2 + the fused LayerNorm, against plain torch on the same inputs."""
    mo = 30
    g0 = torch.Generator().manual_seed(seed)
    boxes = torch.zeros(B, mo, 4)
    masks = torch.zeros(B, mo)
    for b in range(nb):
        n = 5 + b
        xy = torch.rand(n, 2, generator=g0) * 0.6
        wh = torch.rand(n, 2, generator=g0) * 0.35 + 0.05
        boxes[b, :n] = torch.cat([xy, xy + wh], dim=-1)
        masks[b, :n] = 1
    boxes, masks = boxes.to(DEV), masks.to(DEV)
    rects = torch.zeros(B, mo, 5, device=DEV, dtype=torch.int32)
    L.check(L.lib().ltt_op_rela_rects(L.ptr(boxes), L.ptr(masks), B, mo, h, w, L.ptr(rects), L.stream_ptr()), "rects")
    hid = rn(B, h * w, C, seed=seed + 1) * 2
    x = rn(B, h * w, C, seed=seed + 2, dtype=torch.float16)
    feats = rn(nb, mo, C, seed=seed + 3, dtype=torch.float16)
    g = 1 + 0.1 * rn(C, seed=seed + 4)
    bt = 0.1 * rn(C, seed=seed + 5)
    out = torch.empty(B, h * w, C, device=DEV)
    ln16 = torch.empty(B, h * w, C, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_rela_scatter_ln(L.ptr(hid), L.ptr(x), L.ptr(feats), L.ptr(rects), nb, B, mo, h, w, C, L.ptr(out),
                                           L.ptr(g), L.ptr(bt), 1e-5, L.ptr(ln16), L.stream_ptr()), "rela_scatter_ln")
    torch.cuda.synchronize()
    from oracle import unet_oracle as uo
    add = torch.zeros(B, h, w, C, device=DEV)
    for b, row in enumerate(uo.box_pixel_rects(boxes.cpu(), masks.cpu(), h, w)[:nb]):      # the ORACLE's rectangles
        for i, (t, bo, le, ri) in enumerate(row):
            add[b, t:bo, le:ri] += feats[b, i].float() / mo
    ref = ((hid + add.view(B, h * w, C)) + x.float()) * 0.5
    ref_ln = F.layer_norm(ref, (C,), g, bt, 1e-5)
    return max(rel(out, ref), rel(ln16.float(), ref_ln) / 4)


def check_ln_fold(M, K1, C, N, act2=0, seed=0):
    """Producer GEMM with row-statistics epilogue -> consumer GEMM with the LayerNorm folded in, against
    F.linear(F.layer_norm(x)) in fp32 on the same fp16 x (x = the producer's own fp16 output)."""
    a = rn(M, K1, seed=seed, dtype=torch.float16)
    w1 = rn(C, K1, seed=seed + 1, scale=1 / math.sqrt(K1)).half()
    b1 = rn(C, seed=seed + 2, scale=0.5)                      # a row mean well away from 0
    gamma, beta = 1 + 0.2 * rn(C, seed=seed + 3), 0.2 * rn(C, seed=seed + 4)
    w2 = rn(N, C, seed=seed + 5, scale=1 / math.sqrt(C))
    b2 = rn(N, seed=seed + 6, scale=0.1)
    nout = N // 2 if act2 == 2 else N
    out1 = torch.empty(M, C, device=DEV, dtype=torch.float16)
    out2 = torch.empty(M, nout, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_linear_ln_linear(L.ptr(a), M, K1, L.ptr(w1), L.ptr(b1), C, L.ptr(gamma), L.ptr(beta), 1e-5,
                                            L.ptr(w2.contiguous()), L.ptr(b2), N, act2, L.ptr(out1), L.ptr(out2), L.stream_ptr()),
            "linear_ln_linear")
    x = out1.float()
    e1 = rel(x, F.linear(a.float(), w1.float(), b1))
    ref = F.linear(F.layer_norm(x, (C,), gamma, beta, 1e-5), w2.half().float(), b2)
    if act2 == 2:
        v, g = ref.chunk(2, dim=-1)
        ref = v * F.gelu(g)
    return max(e1, rel(out2.float(), ref))

# ------------------------------------------------------------------------------------------------ small ops vs the oracle
def check_rela_rects(B, h, w, seed=0):
    """Integer box rectangles incl. the truncation and `break` rules against oracle.box_pixel_rects (bit-exact)."""
    from oracle import unet_oracle as uo
    mo = 30
    g0 = torch.Generator().manual_seed(seed)
    boxes = torch.zeros(B, mo, 4)
    masks = torch.zeros(B, mo)
    for b in range(B):
        n = [0, 1, 7, 30][b % 4]
        xy = torch.rand(n, 2, generator=g0) * 0.7
        wh = torch.rand(n, 2, generator=g0) * 0.5 + 0.01
        boxes[b, :n] = torch.cat([xy, xy + wh], dim=-1)        # may exceed 1.0: the min(x1*w, w) clamp
        masks[b, :n] = 1
        if n >= 7:
            boxes[b, 3] = torch.tensor([0.5, 0.2, 0.5 + 1e-4, 0.9])   # degenerate (l == r): ends the scan for this sample
    rects = torch.zeros(B, mo, 5, device=DEV, dtype=torch.int32)
    boxes_d, masks_d = boxes.to(DEV), masks.to(DEV)        # named: the device copies must outlive the asynchronous launch
    L.check(L.lib().ltt_op_rela_rects(L.ptr(boxes_d), L.ptr(masks_d), B, mo, h, w, L.ptr(rects), L.stream_ptr()), "rects")
    torch.cuda.synchronize()
    ref = uo.box_pixel_rects(boxes, masks, h, w)
    got = rects.cpu()
    bad = 0
    for b in range(B):
        for i in range(mo):
            ok = int(got[b, i, 4])
            if i < len(ref[b]):
                bad += (ok != 1) or (tuple(int(v) for v in got[b, i, :4]) != tuple(ref[b][i]))
            else:
                bad += ok != 0
    return float(bad)


def check_rela_pool(B, h, w, C, seed=0):
    """Per-box mean pool of the (materialised fp32) LN3 output against torch slicing on the oracle's rectangles."""
    from oracle import unet_oracle as uo
    mo = 30
    g0 = torch.Generator().manual_seed(seed)
    boxes = torch.zeros(B, mo, 4)
    masks = torch.zeros(B, mo)
    for b in range(B):
        n = 3 + 4 * b
        xy = torch.rand(n, 2, generator=g0) * 0.6
        wh = torch.rand(n, 2, generator=g0) * 0.4 + 0.05
        boxes[b, :n] = torch.cat([xy, (xy + wh).clamp(max=1.0)], dim=-1)
        masks[b, :n] = 1
    rects = torch.zeros(B, mo, 5, device=DEV, dtype=torch.int32)
    boxes_d, masks_d = boxes.to(DEV), masks.to(DEV)        # named: the device copies must outlive the asynchronous launch
    L.check(L.lib().ltt_op_rela_rects(L.ptr(boxes_d), L.ptr(masks_d), B, mo, h, w, L.ptr(rects), L.stream_ptr()), "rects")
    hid = rn(B, h * w, C, seed=seed + 1) * 2 + 0.3
    feats = torch.full((B, mo, C), 9.0, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_rela_pool(L.ptr(hid), L.ptr(rects), B, mo, h, w, C, L.ptr(feats), L.stream_ptr()), "rela_pool")
    ref = torch.zeros(B, mo, C, device=DEV)
    hv = hid.view(B, h, w, C)
    for b, row in enumerate(uo.box_pixel_rects(boxes, masks, h, w)):
        for i, (t, bo, le, ri) in enumerate(row):
            ref[b, i] = hv[b, t:bo, le:ri].reshape(-1, C).mean(dim=0)
    return rel(feats.float(), ref)


def check_posnet_input(rows, seed=0):
    """PositionNet input rows [emb*m + (1-m)*null | fourier(box)*m + (1-m)*null] against the oracle (fp16 rounding only)."""
    from oracle import unet_oracle as uo
    boxes = torch.rand(rows, 4, generator=torch.Generator().manual_seed(seed)).to(DEV)
    masks = (torch.arange(rows) % 3 != 2).float().to(DEV)
    emb = rn(rows, 768, seed=seed + 1)
    ntxt, npos = rn(768, seed=seed + 2, scale=0.05), rn(64, seed=seed + 3, scale=0.05)
    out = torch.empty(rows, 832, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_posnet_input(L.ptr(boxes), L.ptr(masks), L.ptr(emb), L.ptr(ntxt), L.ptr(npos), rows, 768, 8,
                                        L.ptr(out), L.stream_ptr()), "posnet_input")
    m = masks[:, None]
    ref = torch.cat([emb * m + (1 - m) * ntxt, uo.fourier_embed(boxes.cpu(), 8).to(DEV) * m + (1 - m) * npos], dim=-1)
    return rel(out.float(), ref.half().float())


def check_timestep_embedding(dim, seed=0):
    """[cos | sin] timestep embedding of every PLMS timestep (1 .. 981) against the oracle, rounded to fp16."""
    from oracle import unet_oracle as uo
    t = torch.arange(1, 1000, 20, dtype=torch.float32, device=DEV)
    out = torch.empty(t.numel(), dim, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_timestep_embedding(L.ptr(t), t.numel(), dim, L.ptr(out), L.stream_ptr()), "timestep_embedding")
    return rel(out.float(), uo.timestep_embedding(t.cpu(), dim).half().float().to(DEV))


def check_plms_update(mode, use_cfg=True, seed=0, n=2 * 4 * 64 * 64):
    """CFG combine + Euler / Adams-Bashforth + x_prev against the reference's expressions (plms.py:116-161) evaluated by
    torch on fp16 eps tensors, which is what those lines compute under autocast (x and the schedule scalars are fp32)."""
    h = lambda s: rn(n, seed=seed + s).half()
    e_c, e_u, o1, o2, o3, ef = h(0), h(1), h(2), h(3), h(4), h(5)
    x = rn(n, seed=seed + 6)
    a_t, a_prev, guidance = 0.41, 0.47, 7.5
    s1m = math.sqrt(1 - a_t)
    c32, u32, f32_, p1, p2, p3 = (t.float().contiguous() for t in (e_c, e_u, ef, o1, o2, o3))    # kept alive across the call
    e_t_out, x_out = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    L.check(L.lib().ltt_op_plms_update(L.ptr(c32), L.ptr(u32), guidance, int(use_cfg), mode, L.ptr(x), L.ptr(e_t_out),
                                       L.ptr(f32_), L.ptr(p1), L.ptr(p2), L.ptr(p3), a_t, a_prev, s1m,
                                       L.ptr(x_out), n, L.stream_ptr()), "plms_update")
    torch.cuda.synchronize()
    e = e_u + guidance * (e_c - e_u) if use_cfg else e_c             # fp16 arithmetic, as under autocast
    if mode == 0:
        ep = e
    elif mode == 1:
        ep = (ef + e) / 2
    elif mode == 2:
        ep = (3 * e - o1) / 2
    elif mode == 3:
        ep = (23 * e - 16 * o1 + 5 * o2) / 12
    else:
        ep = (55 * e - 59 * o1 + 37 * o2 - 9 * o3) / 24
    at, ap, sq = (torch.full((1,), v, device=DEV) for v in (a_t, a_prev, s1m))
    pred_x0 = (x - sq * ep) / at.sqrt()
    ref = ap.sqrt() * pred_x0 + (1. - ap).sqrt() * ep
    err = rel(x_out, ref)
    if mode != 1:
        err = max(err, rel(e_t_out, e.float()))
    return err


ALL = [
    # name, fn, args, tolerance (rel-L2 vs fp32 math on the same fp16 inputs)
    ("linear 256x128x64 (1 tile, 1 k-iter)", check_linear, dict(M=256, N=128, K=64, bias=False), 2e-3),
    ("linear 4096x320x320 bias", check_linear, dict(M=4096, N=320, K=320), 2e-3),
    ("linear 1000x640x1280 tail rows", check_linear, dict(M=1000, N=640, K=1280), 2e-3),
    ("linear 128x1280x5120 split-K", check_linear, dict(M=128, N=1280, K=5120), 2e-3),
    ("linear 60x320x1280 split-K small M", check_linear, dict(M=60, N=320, K=1280), 2e-3),
    ("linear SiLU 90x512x832", check_linear, dict(M=90, N=512, K=832, act=1), 2e-3),
    ("linear GEGLU 1024x2560x320", check_linear, dict(M=1024, N=2560, K=320, act=2), 2e-3),
    ("linear GEGLU split-K 64x10240x1280", check_linear, dict(M=64, N=10240, K=1280, act=2), 2e-3),
    ("linear gate+res16 1024x320x1280", check_linear, dict(M=1024, N=320, K=1280, res_dtype=torch.float16, gate=0.46), 2e-3),
    ("linear res32->f32 512x640x640", check_linear, dict(M=512, N=640, K=640, res_dtype=torch.float32, out_dtype=torch.float32), 2e-3),
    ("linear res32->f16 512x640x2560", check_linear, dict(M=512, N=640, K=2560, res_dtype=torch.float32), 2e-3),
    ("conv3x3 1x64x64 320->320", check_conv3x3, dict(B=1, H=64, W=64, C=320, N=320), 2e-3),
    ("conv3x3 2x32x32 640->640 +rowvec", check_conv3x3, dict(B=2, H=32, W=32, C=640, N=640, rowvec=True), 2e-3),
    ("conv3x3 2x16x16 1280->1280 split-K", check_conv3x3, dict(B=2, H=16, W=16, C=1280, N=1280), 2e-3),
    ("conv3x3 1x8x8 1280->1280 (tb=2 tile)", check_conv3x3, dict(B=1, H=8, W=8, C=1280, N=1280), 2e-3),
    ("conv3x3 3x8x8 64->128", check_conv3x3, dict(B=3, H=8, W=8, C=64, N=128), 2e-3),
    ("conv3x3 2x12x12 128->64 (ragged tile)", check_conv3x3, dict(B=2, H=12, W=12, C=128, N=64), 2e-3),
    ("conv3x3 1x96x96 320->320", check_conv3x3, dict(B=1, H=96, W=96, C=320, N=320), 2e-3),
    ("qkv 2x1024 C=640", check_qkv, dict(B=2, tokens=1024, C=640, heads=8), 2e-3),
    ("qkv 1x4096 C=320", check_qkv, dict(B=1, tokens=4096, C=320, heads=8), 2e-3),
    ("qkv 3x64 C=1280", check_qkv, dict(B=3, tokens=64, C=1280, heads=8), 2e-3),
    ("qkv 2x256 C=64 (tiny heads)", check_qkv, dict(B=2, tokens=256, C=64, heads=8), 2e-3),
    ("attention d=40 nq=256 nk=128 (1 tile)", check_attention, dict(B=1, heads=8, d=40, nq=256, nk=128), 3e-3),
    ("attention d=40 nq=4096 nk=4126", check_attention, dict(B=1, heads=8, d=40, nq=4096, nk=4126), 3e-3),
    ("attention d=80 nq=1024 nk=1054", check_attention, dict(B=2, heads=8, d=80, nq=1024, nk=1054), 3e-3),
    ("attention d=160 nq=256 nk=286", check_attention, dict(B=2, heads=8, d=160, nq=256, nk=286), 3e-3),
    ("attention d=160 nq=64 nk=77", check_attention, dict(B=1, heads=8, d=160, nq=64, nk=77), 3e-3),
    ("attention d=40 nq=4096 nk=77", check_attention, dict(B=1, heads=8, d=40, nq=4096, nk=77), 3e-3),
    ("attention d=40 nq=100 nk=37 (partial q tile, 64-key tail tile)", check_attention, dict(B=2, heads=8, d=40, nq=100, nk=37), 3e-3),
    ("attention d=40 nq=200 nk=129 (128-key tile + 1-key tail)", check_attention, dict(B=1, heads=8, d=40, nq=200, nk=129), 3e-3),
    ("attention d=40 peaky logits (rescale path, 2 warpgroups)", check_attention, dict(B=1, heads=8, d=40, nq=300, nk=700, qscale=6.0), 5e-3),
    ("attention d=80 peaky logits (rescale path)", check_attention, dict(B=1, heads=8, d=80, nq=300, nk=700, qscale=6.0), 5e-3),
    ("attention d=8 nq=256 nk=286", check_attention, dict(B=2, heads=8, d=8, nq=256, nk=286), 3e-3),
    ("attention d=16 nq=64 nk=94", check_attention, dict(B=2, heads=8, d=16, nq=64, nk=94), 3e-3),
    ("groupnorm+silu 2x4096 320", check_groupnorm, dict(B=2, HW=4096, c0=320, c1=0, silu=True, eps=1e-5), 2e-3),
    ("groupnorm concat 1x1024 1280+640", check_groupnorm, dict(B=1, HW=1024, c0=1280, c1=640, silu=True, eps=1e-5), 2e-3),
    ("groupnorm eps1e-6 nosilu 2x64 1280", check_groupnorm, dict(B=2, HW=64, c0=1280, c1=0, silu=False, eps=1e-6), 2e-3),
    ("groupnorm concat 2x4096 320+640 (cpg 30)", check_groupnorm, dict(B=2, HW=4096, c0=320, c1=640, silu=True, eps=1e-5), 2e-3),
    ("groupnorm concat 2x256 1280+1280", check_groupnorm, dict(B=2, HW=256, c0=1280, c1=1280, silu=True, eps=1e-5), 2e-3),
    ("groupnorm 3x4 64 (cpg 2, fewer pixels than CTAs)", check_groupnorm, dict(B=3, HW=4, c0=64, c1=0, silu=True, eps=1e-5), 2e-3),
    ("groupnorm concat 2x384 64+128 (cpg 6, ragged pixels)", check_groupnorm, dict(B=2, HW=381, c0=64, c1=128, silu=False, eps=1e-5), 2e-3),
    ("groupnorm 1x9216 960 (slab not staged)", check_groupnorm, dict(B=1, HW=9216, c0=640, c1=320, silu=True, eps=1e-5), 2e-3),
    ("relation attention folded C=320 10 relations", check_rela_attn_fused, dict(G=2, C=320, heads=8, nrel=10), 2e-3),
    ("relation attention folded C=1280 10 relations", check_rela_attn_fused, dict(G=1, C=1280, heads=8, nrel=10), 2e-3),
    ("relation attention folded C=640 1 relation", check_rela_attn_fused, dict(G=2, C=640, heads=8, nrel=1), 2e-3),
    ("relation attention folded C=64 3 relations", check_rela_attn_fused, dict(G=3, C=64, heads=8, nrel=3), 2e-3),
    ("relation attention 30x10 d=40", check_small_attention, dict(B=2, nq=30, nk=10, heads=8, d=40), 3e-3),
    ("relation attention 30x10 d=160", check_small_attention, dict(B=1, nq=30, nk=10, heads=8, d=160), 3e-3),
    ("relation attention 30x3 d=8", check_small_attention, dict(B=3, nq=30, nk=3, heads=8, d=8), 3e-3),
    ("rela scatter + LN 2x32x32 C=640 (uncond half without boxes)", check_rela_scatter_ln, dict(B=2, nb=1, h=32, w=32, C=640), 1e-4),
    ("rela scatter + LN 3x24x16 C=320", check_rela_scatter_ln, dict(B=3, nb=3, h=24, w=16, C=320), 1e-4),
    ("rela scatter + LN 2x8x8 C=1280", check_rela_scatter_ln, dict(B=2, nb=2, h=8, w=8, C=1280), 1e-4),
    ("rela scatter + LN 2x16x16 C=64", check_rela_scatter_ln, dict(B=2, nb=1, h=16, w=16, C=64), 1e-4),
    ("layernorm f16 4096x320", check_layernorm, dict(M=4096, C=320, dtype=torch.float16), 1e-4),
    ("layernorm f32 1000x1280", check_layernorm, dict(M=1000, C=1280, dtype=torch.float32), 1e-4),
    ("layernorm f16 90x64", check_layernorm, dict(M=90, C=64, dtype=torch.float16), 1e-4),
    # M >= 16384 and C <= 768: the grid-stride variant (1, 2 and 3 vectors per lane; ragged last vector; odd row count)
    ("layernorm rows f16 65536x320", check_layernorm, dict(M=65536, C=320, dtype=torch.float16), 1e-4),
    ("layernorm rows f32 32771x320", check_layernorm, dict(M=32771, C=320, dtype=torch.float32), 1e-4),
    ("layernorm rows f16 16384x640", check_layernorm, dict(M=16384, C=640, dtype=torch.float16), 1e-4),
    ("layernorm rows f16 20001x64", check_layernorm, dict(M=20001, C=64, dtype=torch.float16), 1e-4),
    ("LayerNorm fold 4096x320 -> QKV-like 960", check_ln_fold, dict(M=4096, K1=320, C=320, N=960), 2e-3),
    ("LayerNorm fold 4096x320 -> GEGLU 2560", check_ln_fold, dict(M=4096, K1=320, C=320, N=2560, act2=2), 2e-3),
    ("LayerNorm fold 1000x640 (K1 2560, tail rows) -> 1920", check_ln_fold, dict(M=1000, K1=2560, C=640, N=1920), 2e-3),
    ("LayerNorm fold 128x1280 (split-K producer and consumer) -> 3840", check_ln_fold, dict(M=128, K1=5120, C=1280, N=3840), 2e-3),
    ("LayerNorm fold 60x1280 -> GEGLU 10240 (split-K)", check_ln_fold, dict(M=60, K1=1280, C=1280, N=10240, act2=2), 2e-3),
    ("LayerNorm fold 16384x320 (CTA-pair producer) -> 960", check_ln_fold, dict(M=16384, K1=1280, C=320, N=960), 2e-3),
    ("LayerNorm fold 300x64 -> 192", check_ln_fold, dict(M=300, K1=64, C=64, N=192), 2e-3),
    ("rela rects 8x64x64 vs oracle (truncation, clamp, break rule)", check_rela_rects, dict(B=8, h=64, w=64), 0.5),
    ("rela rects 4x24x16 vs oracle", check_rela_rects, dict(B=4, h=24, w=16, seed=3), 0.5),
    ("rela rects 5x8x8 vs oracle", check_rela_rects, dict(B=5, h=8, w=8, seed=5), 0.5),
    ("rela pool 2x64x64 C=320 vs oracle rectangles", check_rela_pool, dict(B=2, h=64, w=64, C=320), 1e-3),
    ("rela pool 3x8x8 C=1280 vs oracle rectangles", check_rela_pool, dict(B=3, h=8, w=8, C=1280), 1e-3),
    ("PositionNet input rows (Fourier + null mixing) vs oracle", check_posnet_input, dict(rows=90), 1e-3),
    ("timestep embedding dim 320, all PLMS timesteps vs oracle", check_timestep_embedding, dict(dim=320), 1e-3),
    ("timestep embedding dim 64 vs oracle", check_timestep_embedding, dict(dim=64), 1e-3),
    ("plms update mode 0 (Euler predictor) vs reference expressions", check_plms_update, dict(mode=0), 1e-6),
    ("plms update mode 1 (e' = (e_t + e_next)/2)", check_plms_update, dict(mode=1), 1e-6),
    ("plms update mode 2 (AB2)", check_plms_update, dict(mode=2), 1e-6),
    ("plms update mode 3 (AB3)", check_plms_update, dict(mode=3), 1e-6),
    ("plms update mode 4 (AB4)", check_plms_update, dict(mode=4), 1e-6),
    ("plms update mode 4 without CFG", check_plms_update, dict(mode=4, use_cfg=False), 1e-6),
]
