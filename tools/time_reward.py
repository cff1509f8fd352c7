"""Reward-path timing: text features of B captions + image features of B generated and B ground-truth
images + cosine / aesthetic head.  Ours: one text pass, ONE vision pass over the 2B images, one head kernel.  Reference call
pattern on the same GPU (models/policy.py:106-123): eager fp32 transformers CLIPModel.get_text_features, two
get_image_features calls, torch head.  Host-side preprocessing (tokenizer / image processor) is identical for both arms and
left out: both start from ids and pixel_values on the device.  Random-init clip-vit-large-patch14 architecture.
usage: python tools/time_reward.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from layoutllm_t2i_b200.clip import ClipTextEncoder, ClipVisionEncoder  # noqa: E402
from layoutllm_t2i_b200.reward import reward_head  # noqa: E402
from oracle import clip_text_oracle as co  # noqa: E402
from oracle import clip_vision_oracle as cv  # noqa: E402
from oracle.ref_loader import true_fp32  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
DEV = "cuda"
tcfg, vcfg = co.default_clip_text_config(), cv.default_clip_vision_config()
sd = {k: v.to(DEV) for k, v in co.random_state_dict(tcfg, seed=0).items()}
sd.update({k: v.to(DEV) for k, v in cv.random_state_dict(vcfg, seed=1).items()})
aes = {k: v.to(DEV) for k, v in cv.aesthetic_state_dict(768, seed=2).items()}
text, vision = ClipTextEncoder(tcfg, 0), ClipVisionEncoder(vcfg, 0)
text.load_state_dict(sd)
vision.load_state_dict(sd)
text.finalize()
vision.finalize()
ids = co.synthetic_ids(tcfg, [12] * B, L=14, seed=3).to(DEV)
g = torch.Generator().manual_seed(4)
px_pred = (torch.rand(B, 3, 224, 224, generator=g) * 4 - 1.8).to(DEV)
px_gt = (torch.rand(B, 3, 224, 224, generator=g) * 4 - 1.8).to(DEV)
miou, lay = torch.rand(B, generator=g).to(DEV), torch.rand(B, generator=g).to(DEV)


def ours():
    t = text.encode_ids(ids, want_hidden=False, want_embeds=True)[2]
    e = vision.encode(torch.cat([px_pred, px_gt]))[2]
    return reward_head(t, e[:B], e[B:], aes, miou, lay)[0]


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(ours)
n0 = vision.launch_count
vision.encode(torch.cat([px_pred, px_gt]))
flops = 2 * B * (257 * 24 * (4 * 1024 ** 2 + 2 * 1024 * 4096) * 2 + 24 * 4 * 257 * 257 * 1024 + 2 * 256 * 588 * 1024)
print(f"B = {B} samples: sm_100a reward path (text pass + one vision pass over {2 * B} images + head kernel): {ms:.2f} ms "
      f"({vision.launch_count - n0} launches per vision pass; vision tower {flops / 1e12:.2f} TFLOP -> {flops / ms / 1e9:.0f} TFLOP/s incl. the rest)")
# device-side preprocessing of the decoder's uint8 images vs the host processor the reference uses (PIL resize + numpy)
import time  # noqa: E402

import numpy as np  # noqa: E402

imgs = torch.randint(0, 256, (2 * B, 512, 512, 3), dtype=torch.uint8, device=DEV)
pre_ms = timed(lambda: vision.preprocess(imgs))
try:
    from PIL import Image
    from transformers.models.clip import CLIPImageProcessorPil
    proc = CLIPImageProcessorPil()
    host = imgs.cpu().numpy()
    t0 = time.perf_counter()
    pv = proc(images=[Image.fromarray(i) for i in host], return_tensors="pt")["pixel_values"].to(DEV)
    torch.cuda.synchronize()
    host_ms = (time.perf_counter() - t0) * 1e3
    same = bool(torch.equal(pv, vision.preprocess(imgs)))
    print(f"preprocessing of {2 * B} 512x512 images: device kernels {pre_ms:.3f} ms; host CLIPImageProcessor (PIL backend) + H2D {host_ms:.1f} ms "
          f"(+ the D2H of the generated images the reference pays before it); outputs identical: {same}")
except Exception as ex:  # noqa: BLE001
    print(f"preprocessing of {2 * B} 512x512 images: device kernels {pre_ms:.3f} ms (host processor unavailable: {ex!r})")
r_ours = ours()
try:
    from transformers import CLIPConfig, CLIPModel
    tk = ("vocab_size", "max_position_embeddings", "hidden_size", "num_attention_heads", "num_hidden_layers", "intermediate_size",
          "layer_norm_eps", "hidden_act", "eos_token_id")
    vk = ("image_size", "patch_size", "hidden_size", "num_attention_heads", "num_hidden_layers", "intermediate_size", "layer_norm_eps",
          "hidden_act")
    m = CLIPModel(CLIPConfig(text_config=dict({k: tcfg[k] for k in tk}, bos_token_id=tcfg["vocab_size"] - 2, pad_token_id=1),
                             vision_config={k: vcfg[k] for k in vk}, projection_dim=768)).to(DEV).eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k or k == "logit_scale" for k in missing), (missing, unexpected)

    def feats(x):
        return x if torch.is_tensor(x) else x.pooler_output      # transformers 5 returns an output object from get_*_features

    @torch.no_grad()
    def reference():
        t = feats(m.get_text_features(input_ids=ids))
        p = feats(m.get_image_features(pixel_values=px_pred))
        q = feats(m.get_image_features(pixel_values=px_gt))
        return cv.reward_forward(t, p, q, aes, miou, lay)[0]

    ref_ms = timed(reference, reps=5, warm=2)
    with true_fp32():
        r_ref = reference()
    print(f"reference call pattern (eager fp32 transformers {__import__('transformers').__version__}: get_text_features + 2 x get_image_features "
          f"+ torch head): {ref_ms:.1f} ms -> {ref_ms / ms:.1f}x")
    print(f"max |reward - reference reward (TF32 off)| over the batch: {float((r_ours - r_ref).abs().max()):.2e}  (rewards {r_ref.min().item():.2f} .. {r_ref.max().item():.2f})")
except Exception as ex:  # noqa: BLE001
    import traceback
    traceback.print_exc()
    print("reference arm unavailable:", repr(ex)[:200])
