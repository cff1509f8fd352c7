"""Eager fp32 restatement of the CLIP text transformer the reference conditions on.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module: only ``tests/``, the smoke check and
``bench.py`` / ``tools/`` timing legs use it, and only as the checker (or as the timed baseline).

The algorithm is NOT in ``/root/reference``: the reference calls a third-party dependency, Hugging Face ``transformers``
(``CLIPTextModel`` / ``CLIPModel`` of ``openai/clip-vit-large-patch14``; ``requirements.txt`` pins python 3.8 /
pytorch 1.12, ``transformers`` itself is unpinned, the 4.2x line of that time).  Its published algorithm is restated here
on a flat ``state_dict`` in the dependency's own parameter grammar (``text_model.embeddings.token_embedding.weight`` ...),
and parity is anchored on the reference's call sites:

* ``FrozenCLIPEmbedder.forward``   GLIGEN/ldm/modules/encoders/modules.py:157-169 -- ``transformer(input_ids=tokens)`` on
  prompts padded to 77 tokens WITHOUT an attention mask -> ``last_hidden_state`` (UNet context) and ``pooler_output``
  (relation embeddings, txt2img.py:199-209);
* ``encode_one_token``             modules.py:174-182 -- unpadded ids -> ``pooler_output``;
* ``get_clip_feature``             txt2img.py:147-156 -- ``CLIPModel(**inputs)`` on ONE phrase (``padding=True`` of a single
  string pads nothing) -> ``text_model_output.pooler_output``;
* ``extract_text_feat``            txt2img.py:454-457 -- ``CLIPModel.get_text_features`` (pooler_output @ text_projection^T).

Pinning: the installed ``transformers`` (5.5 here) is run on CPU with a seeded random-init ``CLIPTextConfig`` by
``tests/gen_golden_clip.py``; its outputs are committed as ``tests/golden/clip_text.pt`` and checked by
``tests/test_clip_oracle_cpu.py`` (which also compares live when ``transformers`` imports).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def default_clip_text_config() -> dict:
    """Text tower of openai/clip-vit-large-patch14 (the ``version`` of FrozenCLIPEmbedder, modules.py:146)."""
    return dict(vocab_size=49408, max_position_embeddings=77, hidden_size=768, num_attention_heads=12, num_hidden_layers=12,
                intermediate_size=3072, layer_norm_eps=1e-5, hidden_act="quick_gelu", projection_dim=768, eos_token_id=2)


def tiny_clip_text_config() -> dict:
    return dict(vocab_size=1000, max_position_embeddings=77, hidden_size=128, num_attention_heads=2, num_hidden_layers=2,
                intermediate_size=512, layer_norm_eps=1e-5, hidden_act="quick_gelu", projection_dim=64, eos_token_id=2)


def random_state_dict(cfg: dict, seed: int = 0, with_projection: bool = True, outliers: bool = False) -> SD:
    """Seeded random parameters in the dependency's grammar.  ``outliers``: a few embedding channels carry values ~30x the
    rest, as the trained text tower's residual stream does (massive activations), to exercise the fp16 operand range."""
    g = torch.Generator().manual_seed(seed)
    W, Fd, V, P = cfg["hidden_size"], cfg["intermediate_size"], cfg["vocab_size"], cfg["max_position_embeddings"]

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    sd = {"text_model.embeddings.token_embedding.weight": rn(V, W, std=0.3),
          "text_model.embeddings.position_embedding.weight": rn(P, W, std=0.1)}
    if outliers:
        sd["text_model.embeddings.token_embedding.weight"][:, ::97] *= 30.0
    for i in range(cfg["num_hidden_layers"]):
        p = f"text_model.encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"self_attn.{n}.weight"] = rn(W, W, std=1.5 * W ** -0.5)
            sd[p + f"self_attn.{n}.bias"] = rn(W, std=0.1)
        for n in ("layer_norm1", "layer_norm2"):
            sd[p + n + ".weight"] = 1 + rn(W, std=0.2)
            sd[p + n + ".bias"] = rn(W, std=0.1)
        sd[p + "mlp.fc1.weight"] = rn(Fd, W, std=1.5 * W ** -0.5)
        sd[p + "mlp.fc1.bias"] = rn(Fd, std=0.1)
        sd[p + "mlp.fc2.weight"] = rn(W, Fd, std=Fd ** -0.5)
        sd[p + "mlp.fc2.bias"] = rn(W, std=0.1)
    sd["text_model.final_layer_norm.weight"] = 1 + rn(W, std=0.2)
    sd["text_model.final_layer_norm.bias"] = rn(W, std=0.1)
    if with_projection:
        sd["text_projection.weight"] = rn(cfg["projection_dim"], W, std=W ** -0.5)
    return sd


def _act(x: Tensor, name: str) -> Tensor:
    if name == "quick_gelu":                      # transformers activations.QuickGELUActivation
        return x * torch.sigmoid(1.702 * x)
    if name == "gelu":
        return F.gelu(x)
    raise ValueError(f"unsupported hidden_act {name!r}")


def eos_positions(ids: Tensor, eos_token_id: int) -> Tensor:
    """CLIPTextTransformer.forward: the legacy config (eos_token_id == 2, which is what clip-vit-large-patch14 ships) pools
    at argmax(ids) -- the end-of-text token has the largest id -- newer configs at the first occurrence of eos_token_id."""
    if eos_token_id == 2:
        return ids.to(torch.int).argmax(dim=-1)
    return (ids.to(torch.int) == eos_token_id).int().argmax(dim=-1)


def encoder_layers(sd: SD, cfg: dict, x: Tensor, mask: Optional[Tensor], prefix: str) -> Tensor:
    """CLIPEncoder: pre-LayerNorm blocks (CLIPEncoderLayer: x + attn(LN1(x)), x + mlp(LN2(x))) shared by both towers;
    ``mask`` is the additive attention mask ([.., L, L], broadcast over batch / heads) or None (vision tower)."""
    B, L, W = x.shape
    H = cfg["num_attention_heads"]
    d = W // H
    eps = cfg["layer_norm_eps"]
    for i in range(cfg["num_hidden_layers"]):
        p = f"{prefix}encoder.layers.{i}."
        h = F.layer_norm(x, (W,), sd[p + "layer_norm1.weight"], sd[p + "layer_norm1.bias"], eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]) * d ** -0.5
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"])
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
        q, k, v = (t.view(B, L, H, d).transpose(1, 2) for t in (q, k, v))
        s = q @ k.transpose(-1, -2)
        w = torch.softmax(s if mask is None else s + mask, dim=-1)
        a = (w @ v).transpose(1, 2).reshape(B, L, W)
        x = x + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (W,), sd[p + "layer_norm2.weight"], sd[p + "layer_norm2.bias"], eps)
        h = _act(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]), cfg["hidden_act"])
        x = x + F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x


def clip_text_forward(sd: SD, cfg: dict, ids: Tensor, attention_mask: Optional[Tensor] = None,
                      prefix: str = "text_model.") -> Tuple[Tensor, Tensor]:
    """CLIPTextTransformer.forward -> (last_hidden_state [B, L, W], pooler_output [B, W]).

    embeddings (token + learned position) -> pre-LayerNorm blocks with causal self-attention (plus an additive padding
    mask over keys when ``attention_mask`` is given) -> final LayerNorm -> pooled = row of the end-of-text token."""
    B, L = ids.shape
    W = cfg["hidden_size"]
    eps = cfg["layer_norm_eps"]
    x = F.embedding(ids, sd[prefix + "embeddings.token_embedding.weight"]) + sd[prefix + "embeddings.position_embedding.weight"][:L]
    neg = torch.finfo(x.dtype).min
    mask = torch.full((L, L), neg, dtype=x.dtype, device=x.device).triu(1)[None, None]
    if attention_mask is not None:
        mask = mask + (1.0 - attention_mask[:, None, None, :].to(x.dtype)) * neg
        mask = mask.clamp_min(neg)
    x = encoder_layers(sd, cfg, x, mask, prefix)
    x = F.layer_norm(x, (W,), sd[prefix + "final_layer_norm.weight"], sd[prefix + "final_layer_norm.bias"], eps)
    pooled = x[torch.arange(B, device=x.device), eos_positions(ids, cfg["eos_token_id"])]
    return x, pooled


def text_features(sd: SD, cfg: dict, ids: Tensor, attention_mask: Optional[Tensor] = None) -> Tensor:
    """CLIPModel.get_text_features (txt2img.py:454-457): pooler_output through the bias-free text projection."""
    _, pooled = clip_text_forward(sd, cfg, ids, attention_mask)
    return F.linear(pooled, sd["text_projection.weight"])


def synthetic_ids(cfg: dict, lengths, L: Optional[int] = None, seed: int = 0, pad_with_eos: bool = True) -> Tensor:
    """Token rows as CLIPTokenizer lays them out: <bos> words... <eos>, then padding (the tokenizer of
    clip-vit-large-patch14 pads with <eos> = vocab-1; FrozenCLIPEmbedder pads to 77).  ``lengths`` counts the words."""
    g = torch.Generator().manual_seed(seed)
    V = cfg["vocab_size"]
    bos, eos = V - 2, V - 1
    L = cfg["max_position_embeddings"] if L is None else L
    rows = []
    for n in lengths:
        n = min(n, L - 2)
        words = torch.randint(1, V - 2, (n,), generator=g)
        row = torch.cat([torch.tensor([bos]), words, torch.tensor([eos])])
        pad = torch.full((L - row.numel(),), eos if pad_with_eos else 0, dtype=torch.long)
        rows.append(torch.cat([row, pad]))
    return torch.stack(rows)
