"""Pins oracle/clip_text_oracle.py to the third-party implementation the reference calls (Hugging Face `transformers`
CLIPTextModel / CLIPModel text tower; see the oracle's header for the reference call sites): runs the INSTALLED
`transformers` on CPU in fp32 with the oracle's seeded random parameters and stores ids + outputs as
tests/golden/clip_text.pt.  Run here (build container); the fixture travels, `transformers` weights do not exist offline.

usage: python tests/gen_golden_clip.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import clip_text_oracle as co  # noqa: E402


def hf_text_model(cfg, sd):
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel
    keys = ("vocab_size", "max_position_embeddings", "hidden_size", "num_attention_heads", "num_hidden_layers", "intermediate_size",
            "layer_norm_eps", "hidden_act", "projection_dim", "eos_token_id")
    hc = CLIPTextConfig(**{k: cfg[k] for k in keys}, bos_token_id=cfg["vocab_size"] - 2, pad_token_id=1)
    m = CLIPTextModel(hc).eval()
    missing, unexpected = m.load_state_dict({k: v for k, v in sd.items() if k.startswith("text_model.")}, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m, transformers.__version__


def main():
    out = {}
    for name, cfg, seed in (("tiny", co.tiny_clip_text_config(), 3), ("tiny_eos", dict(co.tiny_clip_text_config(), eos_token_id=999), 4)):
        sd = co.random_state_dict(cfg, seed=seed, outliers=True)
        m, ver = hf_text_model(cfg, sd)
        cases = {}
        # (a) FrozenCLIPEmbedder: padded to 77, no attention mask; (b) one unpadded phrase (get_clip_feature /
        # encode_one_token); (c) a padded batch WITH its attention mask (extract_text_feat)
        ids_a = co.synthetic_ids(cfg, [5, 0, 12, 75, 1], seed=seed)
        ids_b = co.synthetic_ids(cfg, [3], L=5, seed=seed + 1)
        ids_c = co.synthetic_ids(cfg, [2, 7, 4], L=9, seed=seed + 2, pad_with_eos=True)
        mask_c = torch.zeros_like(ids_c)
        for r, n in enumerate([2, 7, 4]):
            mask_c[r, :n + 2] = 1
        with torch.no_grad():
            for tag, ids, mask in (("padded77", ids_a, None), ("single", ids_b, None), ("masked", ids_c, mask_c)):
                o = m(input_ids=ids, attention_mask=mask)
                cases[tag] = dict(ids=ids, mask=mask, last_hidden_state=o.last_hidden_state.clone(), pooler_output=o.pooler_output.clone(),
                                  text_embeds=torch.nn.functional.linear(o.pooler_output, sd["text_projection.weight"]))
        out[name] = dict(cfg=cfg, seed=seed, cases=cases, transformers=ver)
        print(name, "transformers", ver, {k: tuple(v["last_hidden_state"].shape) for k, v in cases.items()})
    path = os.path.join(ROOT, "tests", "golden", "clip_text.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


def hf_vision_model(cfg, sd):
    from transformers import CLIPVisionConfig, CLIPVisionModel
    keys = ("image_size", "patch_size", "hidden_size", "num_attention_heads", "num_hidden_layers", "intermediate_size",
            "layer_norm_eps", "hidden_act", "projection_dim")
    m = CLIPVisionModel(CLIPVisionConfig(**{k: cfg[k] for k in keys})).eval()
    missing, unexpected = m.load_state_dict({k: v for k, v in sd.items() if k.startswith("vision_model.")}, strict=False)
    assert not unexpected and all("position_ids" in k for k in missing), (missing, unexpected)
    return m


def main_vision():
    from oracle import clip_vision_oracle as cv
    import transformers
    cfg = cv.tiny_clip_vision_config()
    sd = cv.random_state_dict(cfg, seed=6)
    m = hf_vision_model(cfg, sd)
    px = torch.randn(3, 3, cfg["image_size"], cfg["image_size"], generator=torch.Generator().manual_seed(1)) * 1.2
    with torch.no_grad():
        o = m(pixel_values=px)
    out = dict(cfg=cfg, seed=6, pixel_values=px, last_hidden_state=o.last_hidden_state.clone(), pooler_output=o.pooler_output.clone(),
               image_embeds=torch.nn.functional.linear(o.pooler_output, sd["visual_projection.weight"]), transformers=transformers.__version__)
    path = os.path.join(ROOT, "tests", "golden", "clip_vision.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes", tuple(o.last_hidden_state.shape))


def main_preprocess():
    """Outputs of the PIL-backed transformers CLIP image processor (the reference-era behaviour; the default processor of
    transformers 5 resizes with torchvision and differs by one grey level on ~0.4 % of the pixels) on small seeded images."""
    import numpy as np
    from PIL import Image
    from transformers.models.clip import CLIPImageProcessorPil
    rng = np.random.default_rng(5)
    shapes = [(70, 90), (64, 64), (33, 120), (32, 32), (20, 27)]          # down- and up-scaling, crop on either axis, identity
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    p = CLIPImageProcessorPil(size={"shortest_edge": 32}, crop_size={"height": 32, "width": 32})
    pv = np.concatenate([p(images=[Image.fromarray(i)], return_tensors="np")["pixel_values"] for i in imgs])
    path = os.path.join(ROOT, "tests", "golden", "clip_preprocess.npz")
    np.savez_compressed(path, pixel_values=pv, size=32, **{f"img{i}": im for i, im in enumerate(imgs)})
    print("wrote", path, os.path.getsize(path), "bytes", pv.shape)


if __name__ == "__main__":
    main_preprocess()
    main_vision()
    main()
