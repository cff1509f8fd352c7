"""Conditioning-prep parity: the sm_100a CLIP text tower through the C-ABI
(layoutllm_t2i_b200.clip.ClipTextEncoder / the drop-in FrozenCLIPEmbedder / prepare_conditioning) against
oracle/clip_text_oracle.py -- the fp32 restatement pinned to the transformers implementation the reference calls -- on the
same seeded weights and token ids.  The reference computes in fp32; the engine uses fp16 tensor-core operands with fp32
accumulation and an fp32 residual stream, so the gate is a relative-L2 tolerance, written with each case.
Used by tests/test_clip_gpu.py."""
from __future__ import annotations

import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutllm_t2i_b200 import _lib as L  # noqa: E402
from oracle import clip_text_oracle as co  # noqa: E402
from oracle.ref_loader import true_fp32  # noqa: E402

DEV = "cuda"
_CACHE = {}
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs"))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rn(*shape, seed=0, dtype=torch.float32):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)).to(DEV).to(dtype)


# ------------------------------------------------------------------------------------------------- operator level
def check_causal_attention(B, heads, d, dpad, n, seed=0):
    """ltt_op_attention_causal against fp32 torch math on the same fp16 q / k / v."""
    C = heads * d
    q = rn(B, n, C, seed=seed, dtype=torch.float16)
    k = rn(B, n, C, seed=seed + 1, dtype=torch.float16)
    v = rn(B, n, C, seed=seed + 2, dtype=torch.float16)
    qp = torch.zeros(B, n, heads * dpad, device=DEV, dtype=torch.float16)
    kp = torch.zeros_like(qp)
    qp.view(B, n, heads, dpad)[..., :d] = q.view(B, n, heads, d)
    kp.view(B, n, heads, dpad)[..., :d] = k.view(B, n, heads, d)
    pitch = (n + 7) // 8 * 8
    vt = torch.zeros(B, C, pitch, device=DEV, dtype=torch.float16)
    vt[:, :, :n] = v.transpose(1, 2)
    out = torch.empty(B, n, C, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_attention_causal(L.ptr(qp), n, L.ptr(kp), n, L.ptr(vt), pitch, B, heads, d, dpad, n, d ** -0.5,
                                            L.ptr(out), C, L.stream_ptr()), "attention_causal")
    qf, kf, vf = (t.float().view(B, n, heads, d).transpose(1, 2) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * d ** -0.5
    s = s.masked_fill(torch.ones(n, n, device=DEV, dtype=torch.bool).triu(1), float("-inf"))
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, n, C)
    return rel(out.float(), ref)


def check_linear_quick_gelu(M, N, K, seed=0):
    a = rn(M, K, seed=seed, dtype=torch.float16)
    w = (rn(N, K, seed=seed + 1) * K ** -0.5).half()
    b = rn(N, seed=seed + 2) * 0.1
    out = torch.empty(M, N, device=DEV, dtype=torch.float16)
    L.check(L.lib().ltt_op_linear(L.ptr(a), M, K, K, L.ptr(w), N, L.ptr(b), 3, None, 0, 0, 1.0, 0, L.ptr(out), 0, N,
                                  L.stream_ptr()), "linear quick_gelu")
    y = a.float() @ w.float().t() + b
    return rel(out.float(), y * torch.sigmoid(1.702 * y))


# ------------------------------------------------------------------------------------------------- text tower
def _cfg(which):
    if which == "small768":      # few output tiles and a deep K at one short string: the shape where the GEMM launcher prefers split-K
        return dict(co.default_clip_text_config(), vocab_size=1000, num_hidden_layers=2, intermediate_size=1024)
    return co.tiny_clip_text_config() if which == "tiny" else co.default_clip_text_config()


def tower(which, seed, outliers=False):
    key = (which, seed, outliers)
    if key not in _CACHE:
        from layoutllm_t2i_b200.clip import ClipTextEncoder
        cfg = _cfg(which)
        sd = {k: v.to(DEV) for k, v in co.random_state_dict(cfg, seed=seed, outliers=outliers).items()}
        enc = ClipTextEncoder(cfg, 0)
        enc.load_state_dict(sd)
        enc.finalize()
        _CACHE[key] = (cfg, sd, enc)
    return _CACHE[key]


def oracle(cfg, sd, ids, mask=None):
    with torch.no_grad(), true_fp32():
        return co.clip_text_forward(sd, cfg, ids.to(DEV), None if mask is None else mask.to(DEV))


def check_tower(which, lengths, L_=None, seed=0, outliers=False, what="hidden"):
    cfg, sd, enc = tower(which, seed, outliers)
    ids = co.synthetic_ids(cfg, lengths, L=L_, seed=seed + 17)
    hid, pooled, emb = enc.encode_ids(ids, want_embeds=True)
    z, p = oracle(cfg, sd, ids)
    if what == "hidden":
        return rel(hid, z)
    if what == "pooled":
        return rel(pooled, p)
    with true_fp32():
        return rel(emb, torch.nn.functional.linear(p, sd["text_projection.weight"]))


def check_tower_golden(name):
    """Engine against the committed transformers outputs themselves (tests/golden/clip_text.pt)."""
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clip_text.pt"))[name]
    from layoutllm_t2i_b200.clip import ClipTextEncoder
    enc = ClipTextEncoder(g["cfg"], 0)
    enc.load_state_dict(co.random_state_dict(g["cfg"], seed=g["seed"], outliers=True))
    worst = 0.0
    for tag, c in g["cases"].items():
        hid, pooled, emb = enc.encode_ids(c["ids"], want_embeds=True)
        if c["mask"] is None:      # with a padding mask only the rows up to <eos> (and the pooled vector) are comparable
            worst = max(worst, rel(hid, c["last_hidden_state"]))
        worst = max(worst, rel(pooled, c["pooler_output"]), rel(emb, c["text_embeds"]))
    enc.close()
    return worst


def check_pooled_matches_per_phrase(which, seed=0):
    """One padded batch against the reference's call pattern: every phrase alone, unpadded (get_clip_feature,
    encode_one_token) -- the pooled vectors must agree to the same tolerance as a like-for-like comparison."""
    cfg, sd, enc = tower(which, seed)
    words = [1, 2, 3, 5, 9, 30, 75]
    ids = co.synthetic_ids(cfg, words, seed=seed + 5)
    _, pooled = enc.encode_ids(ids, want_hidden=False)
    ref = torch.cat([oracle(cfg, sd, ids[r:r + 1, :n + 2])[1] for r, n in enumerate(words)])
    one = torch.cat([enc.encode_ids(ids[r:r + 1, :n + 2])[1] for r, n in enumerate(words)])
    return max(rel(pooled, ref), rel(one, ref))


def check_deterministic(which, seed=0):
    cfg, sd, enc = tower(which, seed)
    ids = co.synthetic_ids(cfg, [4, 7, 11], seed=3)
    a = [t.clone() for t in enc.encode_ids(ids)]
    b = enc.encode_ids(ids)
    return 0.0 if all(torch.equal(x, y) for x, y in zip(a, b)) else 1.0


# ------------------------------------------------------------------------------------------------- host logic
def _reference_prep(cfg, sd, tok, prompt, phrases, locations, relations, batch, max_objs=30, max_relas=5):
    """The reference's sequence of calls (txt2img.py:173-209, 213-244, 268-277) with the oracle as the text tower."""
    def enc77(texts):
        ids = tok(texts, truncation=True, max_length=77, padding="max_length")["input_ids"]
        return oracle(cfg, sd, ids)
    W = cfg["hidden_size"]
    boxes, masks, tmask, temb = torch.zeros(max_objs, 4), torch.zeros(max_objs), torch.zeros(max_objs), torch.zeros(max_objs, W)
    feats = [None if p is None else oracle(cfg, sd, tok(p, padding=True)["input_ids"])[1].cpu() for p in phrases]
    for idx, (box, f) in enumerate(zip(locations, feats)):
        boxes[idx] = torch.tensor(box)
        masks[idx] = 1
        if f is not None:
            temb[idx] = f
            tmask[idx] = 1
    context = enc77([prompt] * batch)[0]
    rel_emb = torch.zeros(max_relas, W)
    if relations:
        rel_emb[:len(relations)] = enc77(relations[:max_relas])[1].cpu()
    uc = enc77(batch * [""])[0]
    rep = lambda t: t.unsqueeze(0).repeat(batch, *([1] * t.dim()))  # noqa: E731
    return dict(context=context, uc=uc, relations=rep(rel_emb), boxes=rep(boxes), masks=rep(masks), text_masks=rep(tmask),
                text_embeddings=rep(temb))


def check_prepare_conditioning(which, n_boxes, n_rel, batch=2, seed=0, with_none=False):
    from ltt_test_stubs import HashTokenizer
    from layoutllm_t2i_b200.clip import prepare_conditioning, relation_phrases
    cfg, sd, enc = tower(which, seed)
    tok = HashTokenizer(cfg["vocab_size"])
    prompt = "a cat sitting on a wooden bench next to a red bicycle in the park"
    names = ["cat", "wooden bench", "red bicycle", "park", "tree", "dog on a leash"]
    phrases = [names[i % len(names)] + (" %d" % i if i >= len(names) else "") for i in range(n_boxes)]
    if with_none and phrases:
        phrases[len(phrases) // 2] = None
    g = torch.Generator().manual_seed(seed)
    locations = [sorted(torch.rand(2, generator=g).tolist()) + sorted(torch.rand(2, generator=g).tolist()) for _ in range(n_boxes)]
    graph = dict(entities=[dict(lemma_head="cat"), dict(lemma_head="bench"), dict(lemma_head="bicycle")],
                 relations=[dict(subject=0, relation="sitting on", object=1), dict(subject=1, relation="next to", object=2)][:n_rel])
    relations = relation_phrases(graph, 5)
    out = prepare_conditioning(enc, lambda t: tok(t, truncation=True, max_length=77, padding="max_length")["input_ids"],
                               prompt, phrases, locations, relations, batch=batch)
    ref = _reference_prep(cfg, sd, tok, prompt, phrases, locations, relations, batch)
    worst = 0.0
    for k, v in ref.items():
        o = out[k]
        assert o.shape == v.shape, (k, o.shape, v.shape)
        if k in ("boxes", "masks", "text_masks"):
            assert torch.equal(o.cpu(), v.cpu()), k
        elif float(v.abs().max()) == 0.0:
            assert float(o.abs().max()) == 0.0, k
        else:
            worst = max(worst, rel(o, v))
    return worst


def check_relation_phrase_list():
    """prepare_relation_phrases' list (txt2img.py:218-238): PAD + triplets twice, cut to max_relas; none -> nothing encoded."""
    from layoutllm_t2i_b200.clip import relation_phrases
    ent = [dict(lemma_head=w) for w in ("cat", "bench", "bike")]
    r1 = [dict(subject=0, relation="on", object=1)]
    r3 = r1 + [dict(subject=1, relation="near", object=2), dict(subject=0, relation="behind", object=2)]
    ok = relation_phrases(dict(entities=ent, relations=r1)) == ["PAD", "cat on bench", "cat on bench"]
    ok &= relation_phrases(dict(entities=ent, relations=r3)) == ["PAD", "cat on bench", "bench near bike", "cat behind bike", "cat on bench"]
    ok &= relation_phrases(dict(entities=ent, relations=[])) == [] and relation_phrases(dict(entities=ent)) == []
    return 0.0 if ok else 1.0


def check_dropin_embedder(seed=0):
    """The drop-in FrozenCLIPEmbedder: reference constructor + strict load_state_dict in transformers' grammar, then
    encode / encode(return_pooler_output) / encode_one_token against the oracle."""
    from ltt_test_stubs import HashTokenizer
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "layoutllm_t2i_b200", "dropin")
    if d not in sys.path:
        sys.path.insert(0, d)
    import _ltt_dropin_hook
    _ltt_dropin_hook.install()
    from ldm.util import instantiate_from_config
    cfg = co.tiny_clip_text_config()
    sd = co.random_state_dict(cfg, seed=seed, with_projection=False)
    m = instantiate_from_config(dict(target="ldm.modules.encoders.modules.FrozenCLIPEmbedder",
                                     params=dict(text_config={k: v for k, v in cfg.items() if k != "projection_dim"})))
    ck = {"transformer." + k: v for k, v in sd.items()}
    ck["transformer.text_model.embeddings.position_ids"] = torch.arange(77)[None]
    m.load_state_dict(ck)                      # strict, as txt2img.py:108
    m = m.to(DEV).eval()
    if "device" in vars(m):                    # txt2img.py:112-114
        m.device = DEV
    m.set_tokenizer(HashTokenizer(cfg["vocab_size"]))
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    texts = ["a photo of a cat", "", "two dogs playing with a ball on the beach"]
    ids = m.tokenize_padded(texts)
    z_ref, p_ref = oracle(cfg, sdd, ids)
    z = m.encode(texts)
    z2, p = m.encode(texts, return_pooler_output=True)
    one = m.encode_one_token("wooden bench")
    one_ref = oracle(cfg, sdd, m.tokenizer(text="wooden bench")["input_ids"])[1]
    assert torch.equal(z, z2) and z.shape == (3, 77, cfg["hidden_size"]) and one.shape == (1, cfg["hidden_size"])
    return max(rel(z, z_ref), rel(p, p_ref), rel(one, one_ref))


def check_dropin_vs_reference_class(which="tiny", seed=0):
    """The drop-in FrozenCLIPEmbedder against the reference's OWN class (GLIGEN/ldm/modules/encoders/modules.py:144-182,
    byte-identical copy under oracle/_ref) running transformers' CLIPTextModel in eager fp32 on the same GPU with the same
    weights: encode, encode(return_pooler_output=True) and encode_one_token.  The reference class is constructed without
    `from_pretrained` (no network); tokenizer = the same stand-in on both sides."""
    from ltt_test_stubs import HashTokenizer
    from transformers import CLIPTextConfig, CLIPTextModel
    from oracle import ref_loader as rl
    assert rl.available(), "oracle/_ref is not staged"
    with rl.reference_tree():
        from ldm.modules.encoders.modules import FrozenCLIPEmbedder as RefEmbedder
    cfg = _cfg(which)
    sd = co.random_state_dict(cfg, seed=seed, with_projection=False)
    keys = ("vocab_size", "max_position_embeddings", "hidden_size", "num_attention_heads", "num_hidden_layers", "intermediate_size",
            "layer_norm_eps", "hidden_act", "projection_dim", "eos_token_id")
    hf = CLIPTextModel(CLIPTextConfig(**{k: cfg[k] for k in keys}, bos_token_id=cfg["vocab_size"] - 2, pad_token_id=1)).to(DEV).eval()
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    tok = HashTokenizer(cfg["vocab_size"])

    class Tok:      # CLIPTokenizer returns a BatchEncoding (dict access) -- what the reference's forward indexes
        def __call__(self, text=None, **kw):
            return tok(text, **kw)
    ref = object.__new__(RefEmbedder)
    torch.nn.Module.__init__(ref)
    ref.tokenizer, ref.transformer, ref.device, ref.max_length = Tok(), hf, DEV, 77
    # ---- ours, through the reference's config-string instantiation
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "layoutllm_t2i_b200", "dropin")
    if d not in sys.path:
        sys.path.insert(0, d)
    import _ltt_dropin_hook
    _ltt_dropin_hook.install()
    from ldm.util import instantiate_from_config
    m = instantiate_from_config(dict(target="ldm.modules.encoders.modules.FrozenCLIPEmbedder",
                                     params=dict(text_config={k: v for k, v in cfg.items() if k != "projection_dim"})))
    m.load_state_dict(ref.state_dict())            # the reference module's own state_dict, strict
    m = m.to(DEV).eval()
    m.device = DEV
    m.set_tokenizer(Tok())
    texts = ["a photo of a cat", "", "two dogs playing with a ball on the beach"]
    with torch.no_grad(), true_fp32():
        z_ref = ref.encode(texts)
        z2_ref, p_ref = ref.encode(texts, return_pooler_output=True)
        one_ref = ref.encode_one_token("wooden bench")
    z, (z2, p), one = m.encode(texts), m.encode(texts, return_pooler_output=True), m.encode_one_token("wooden bench")
    assert z.shape == z_ref.shape and p.shape == p_ref.shape and one.shape == one_ref.shape and torch.equal(z_ref, z2_ref)
    return max(rel(z, z_ref), rel(z2, z_ref), rel(p, p_ref), rel(one, one_ref))


# fp16 operands / fp32 accumulate against an fp32 reference: 2e-3 relative L2 (measured values in profiles/r02_clip_parity.txt)
TOL = 2e-3
ALL = [
    ("causal attention d=64 77 tokens", check_causal_attention, dict(B=3, heads=12, d=64, dpad=64, n=77), 1e-3),
    ("causal attention d=64 1 token", check_causal_attention, dict(B=2, heads=2, d=64, dpad=64, n=1), 1e-3),
    ("causal attention d=64 300 tokens (3 query tiles)", check_causal_attention, dict(B=2, heads=3, d=64, dpad=64, n=300), 1e-3),
    ("causal attention d=64 129 tokens", check_causal_attention, dict(B=1, heads=4, d=64, dpad=64, n=129), 1e-3),
    ("linear quick_gelu 154x3072x768", check_linear_quick_gelu, dict(M=154, N=3072, K=768), 1e-3),
    ("relation phrase list", check_relation_phrase_list, {}, 0.5),
    ("tower tiny vs transformers fixture", check_tower_golden, dict(name="tiny"), TOL),
    ("tower tiny (first-eos pooling) vs transformers fixture", check_tower_golden, dict(name="tiny_eos"), TOL),
    ("tower tiny hidden B=5 L=77", check_tower, dict(which="tiny", lengths=[5, 0, 12, 75, 1]), TOL),
    ("tower tiny pooled L=9", check_tower, dict(which="tiny", lengths=[3, 7, 1], L_=9, what="pooled"), TOL),
    ("tower tiny outlier channels", check_tower, dict(which="tiny", lengths=[8, 20], seed=5, outliers=True), TOL),
    ("tower 768 / ffn 1024, one string (few-tile GEMMs)", check_tower, dict(which="small768", lengths=[6]), TOL),
    ("tower 768 / ffn 1024, one unpadded phrase L=5", check_tower, dict(which="small768", lengths=[3], L_=5, what="pooled"), TOL),
    ("tower ViT-L/14 hidden B=1 L=77", check_tower, dict(which="full", lengths=[9]), TOL),
    ("tower ViT-L/14 hidden B=2 L=77", check_tower, dict(which="full", lengths=[12, 0]), TOL),
    ("tower ViT-L/14 pooled B=37 L=77", check_tower, dict(which="full", lengths=[14, 0] + [2] * 30 + [5] * 5, what="pooled"), TOL),
    ("tower ViT-L/14 text_embeds B=4 L=20", check_tower, dict(which="full", lengths=[3, 18, 9, 1], L_=20, what="embeds"), TOL),
    ("tower ViT-L/14 outlier channels", check_tower, dict(which="full", lengths=[10, 40, 75], seed=2, outliers=True), TOL),
    ("batched pass == per-phrase calls (tiny)", check_pooled_matches_per_phrase, dict(which="tiny"), TOL),
    ("batched pass == per-phrase calls (ViT-L/14)", check_pooled_matches_per_phrase, dict(which="full"), TOL),
    ("tower deterministic", check_deterministic, dict(which="full"), 0.5),
    ("prepare_conditioning 6 boxes 2 relations", check_prepare_conditioning, dict(which="full", n_boxes=6, n_rel=2), TOL),
    ("prepare_conditioning 30 boxes 1 relation, a None phrase", check_prepare_conditioning,
     dict(which="full", n_boxes=30, n_rel=1, with_none=True), TOL),
    ("prepare_conditioning 0 boxes 0 relations", check_prepare_conditioning, dict(which="tiny", n_boxes=0, n_rel=0, batch=1), TOL),
    ("drop-in FrozenCLIPEmbedder", check_dropin_embedder, {}, TOL),
    ("drop-in FrozenCLIPEmbedder vs the reference's own class (eager fp32 transformers), tiny", check_dropin_vs_reference_class, {}, TOL),
    ("drop-in FrozenCLIPEmbedder vs the reference's own class, ViT-L/14", check_dropin_vs_reference_class, dict(which="full", seed=4), TOL),
]


def main():
    for name, fn, kw, tol in ALL:
        try:
            err = fn(**kw)
            torch.cuda.synchronize()
        except Exception as ex:  # noqa: BLE001
            print(f"EXC  {name:64s} {ex!r}"[:300], flush=True)
            continue
        print(f"{'ok ' if err < tol else 'FAIL'} {name:64s} {err:.3e}  (gate {tol:g})", flush=True)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    main()
