"""Eager fp32 restatement of the CLIP vision tower and of the reward head built on it.

TEST INFRASTRUCTURE ONLY (same rule as the other oracle modules: imported by ``tests/``, the smoke check and timing tools
only, never by the product package).

Reference call sites (/root/reference): ``Reward.forward`` models/policy.py:106-123 -- ``CLIPModel.get_text_features`` on
the captions, ``CLIPModel.get_image_features`` on the generated and the ground-truth images, cosine similarities,
``AestheticMLP`` (tools/aesthetic.py:9-31) on the L2-normalised image embedding (``normalized``, :52-57) -- and :139
(``clip_reward + 0.1 * aes + 10 * miou + 10 * laysim``; the two layout terms are host-side numpy / scipy on a handful of
boxes, tools/metrics.py, and are passed in).  The vision tower itself is third-party: transformers
``CLIPVisionTransformer`` of ``openai/clip-vit-large-patch14``; restated here in its parameter grammar and pinned to the
installed ``transformers`` by ``tests/gen_golden_clip.py`` -> ``tests/golden/clip_vision.pt``
(``tests/test_clip_oracle_cpu.py``).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from .clip_text_oracle import encoder_layers

Tensor = torch.Tensor
SD = Dict[str, Tensor]


def default_clip_vision_config() -> dict:
    """Vision tower of openai/clip-vit-large-patch14."""
    return dict(image_size=224, patch_size=14, hidden_size=1024, num_attention_heads=16, num_hidden_layers=24,
                intermediate_size=4096, layer_norm_eps=1e-5, hidden_act="quick_gelu", projection_dim=768)


def tiny_clip_vision_config() -> dict:
    return dict(image_size=56, patch_size=14, hidden_size=128, num_attention_heads=2, num_hidden_layers=2,
                intermediate_size=512, layer_norm_eps=1e-5, hidden_act="quick_gelu", projection_dim=64)


def random_state_dict(cfg: dict, seed: int = 0) -> SD:
    g = torch.Generator().manual_seed(seed)
    W, Fd, P = cfg["hidden_size"], cfg["intermediate_size"], cfg["patch_size"]
    n_pos = (cfg["image_size"] // P) ** 2 + 1

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    v = "vision_model."
    sd = {v + "embeddings.class_embedding": rn(W, std=0.3),
          v + "embeddings.patch_embedding.weight": rn(W, 3, P, P, std=(3 * P * P) ** -0.5),
          v + "embeddings.position_embedding.weight": rn(n_pos, W, std=0.1),
          v + "pre_layrnorm.weight": 1 + rn(W, std=0.2), v + "pre_layrnorm.bias": rn(W, std=0.1),
          v + "post_layernorm.weight": 1 + rn(W, std=0.2), v + "post_layernorm.bias": rn(W, std=0.1),
          "visual_projection.weight": rn(cfg["projection_dim"], W, std=W ** -0.5)}
    for i in range(cfg["num_hidden_layers"]):
        p = f"{v}encoder.layers.{i}."
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
    return sd


def clip_vision_forward(sd: SD, cfg: dict, pixel_values: Tensor, prefix: str = "vision_model.") -> Tuple[Tensor, Tensor]:
    """CLIPVisionTransformer.forward -> (last_hidden_state [B, 1 + n_patches, W], pooler_output [B, W]).

    CLIPVisionEmbeddings (stride-P patch convolution without bias, class token in front, learned positions) ->
    pre_layrnorm (sic) -> encoder without a mask -> pooled = post_layernorm(class-token row); last_hidden_state is
    returned WITHOUT post_layernorm."""
    W, eps = cfg["hidden_size"], cfg["layer_norm_eps"]
    B = pixel_values.shape[0]
    pe = F.conv2d(pixel_values, sd[prefix + "embeddings.patch_embedding.weight"], stride=cfg["patch_size"])
    pe = pe.flatten(2).transpose(1, 2)
    x = torch.cat([sd[prefix + "embeddings.class_embedding"].expand(B, 1, -1), pe], dim=1)
    x = x + sd[prefix + "embeddings.position_embedding.weight"]
    x = F.layer_norm(x, (W,), sd[prefix + "pre_layrnorm.weight"], sd[prefix + "pre_layrnorm.bias"], eps)
    x = encoder_layers(sd, cfg, x, None, prefix)
    pooled = F.layer_norm(x[:, 0, :], (W,), sd[prefix + "post_layernorm.weight"], sd[prefix + "post_layernorm.bias"], eps)
    return x, pooled


def image_features(sd: SD, cfg: dict, pixel_values: Tensor) -> Tensor:
    """CLIPModel.get_image_features (models/policy.py:111-114): pooler_output through the bias-free visual projection."""
    return F.linear(clip_vision_forward(sd, cfg, pixel_values)[1], sd["visual_projection.weight"])


# ------------------------------------------------------------------------------------------------- reward head
def aesthetic_state_dict(input_size: int, seed: int = 0) -> SD:
    """AestheticMLP.layers (tools/aesthetic.py:15-27): Linear 1024 / 128 / 64 / 16 / 1 with dropouts between (identity in
    eval) and NO activation."""
    g = torch.Generator().manual_seed(seed)
    sd, dims = {}, [input_size, 1024, 128, 64, 16, 1]
    for idx, (i, o) in zip((0, 2, 4, 6, 7), zip(dims[:-1], dims[1:])):
        sd[f"layers.{idx}.weight"] = torch.randn(o, i, generator=g) * i ** -0.5
        sd[f"layers.{idx}.bias"] = torch.randn(o, generator=g) * 0.1
    return sd


def aesthetic_forward(sd: SD, x: Tensor) -> Tensor:
    for idx in (0, 2, 4, 6, 7):
        x = F.linear(x, sd[f"layers.{idx}.weight"], sd[f"layers.{idx}.bias"])
    return x


def reward_forward(txt_features: Tensor, pred_features: Tensor, gt_features: Tensor, aes_sd: SD, miou: Tensor,
                   laysim: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """Reward.forward models/policy.py:115-139 from the three feature matrices -> (reward, clip_reward, aes_reward)."""
    t, p, g = (F.normalize(x, dim=-1) for x in (txt_features, pred_features, gt_features))
    clip_reward = (t * p).sum(dim=-1) + (g * p).sum(dim=-1)
    n = p.norm(dim=-1, keepdim=True)          # `normalized` (tools/aesthetic.py:52-57) applied to the already normalised features
    n = torch.where(n == 0, torch.ones_like(n), n)
    aes = aesthetic_forward(aes_sd, p / n).flatten()
    return clip_reward + aes * 0.1 + miou * 10 + laysim * 10, clip_reward, aes
