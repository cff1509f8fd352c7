"""Reward-path parity: the sm_100a CLIP vision tower, the reward-head kernel and the Reward mirror
(layoutllm_t2i_b200.reward) through the C-ABI against oracle/clip_vision_oracle.py -- the fp32 restatement pinned to the
transformers implementation / the reference's own AestheticMLP -- on the same seeded weights and inputs.  The reference
computes the towers in fp32; the engine uses fp16 tensor-core operands with fp32 accumulation, so tower gates are relative-L2
tolerances (written with each case); the reward head itself is fp32 arithmetic (1e-5).
Used by tests/test_reward_gpu.py."""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs"))
from oracle import clip_text_oracle as co  # noqa: E402
from oracle import clip_vision_oracle as cv  # noqa: E402
from oracle.ref_loader import true_fp32  # noqa: E402

DEV = "cuda"
_CACHE = {}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _vcfg(which):
    if which == "tiny224":
        return dict(cv.tiny_clip_vision_config(), image_size=224)
    return cv.tiny_clip_vision_config() if which == "tiny" else cv.default_clip_vision_config()


def vision(which, seed):
    key = ("v", which, seed)
    if key not in _CACHE:
        from layoutllm_t2i_b200.clip import ClipVisionEncoder
        cfg = _vcfg(which)
        sd = {k: v.to(DEV) for k, v in cv.random_state_dict(cfg, seed=seed).items()}
        enc = ClipVisionEncoder(cfg, 0)
        enc.load_state_dict(sd)
        enc.finalize()
        _CACHE[key] = (cfg, sd, enc)
    return _CACHE[key]


def pixels(cfg, B, seed=0):
    # CLIPProcessor output range: (x / 255 - mean) / std  ~  [-1.8, 2.2]
    return (torch.rand(B, 3, cfg["image_size"], cfg["image_size"], generator=torch.Generator().manual_seed(seed)) * 4.0 - 1.8).to(DEV)


def check_vision(which, B, seed=0, what="hidden"):
    cfg, sd, enc = vision(which, seed)
    px = pixels(cfg, B, seed + 3)
    hid, pooled, emb = enc.encode(px, want_hidden=True)
    with torch.no_grad(), true_fp32():
        z, p = cv.clip_vision_forward(sd, cfg, px)
        e = torch.nn.functional.linear(p, sd["visual_projection.weight"])
    return {"hidden": rel(hid, z), "pooled": rel(pooled, p), "embeds": rel(emb, e)}[what]


def check_vision_golden():
    """Engine against the committed transformers outputs themselves (tests/golden/clip_vision.pt)."""
    from layoutllm_t2i_b200.clip import ClipVisionEncoder
    g = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clip_vision.pt"))
    enc = ClipVisionEncoder(g["cfg"], 0)
    enc.load_state_dict(cv.random_state_dict(g["cfg"], seed=g["seed"]))
    hid, pooled, emb = enc.encode(g["pixel_values"], want_hidden=True)
    err = max(rel(hid, g["last_hidden_state"]), rel(pooled, g["pooler_output"]), rel(emb, g["image_embeds"]))
    enc.close()
    return err


def check_vision_batch_invariance(which="tiny"):
    """image i of a batch == image i alone (no cross-image leakage through the [B * T] row packing), to fp16 tile noise."""
    cfg, sd, enc = vision(which, 0)
    px = pixels(cfg, 4, 11)
    all_ = enc.encode(px)[2].clone()
    one = torch.cat([enc.encode(px[i:i + 1])[2] for i in range(4)])
    return rel(one, all_)


def check_reward_head(B, D, seed=0, with_layout=True, zero_row=False):
    from layoutllm_t2i_b200.reward import reward_head
    g = torch.Generator().manual_seed(seed)
    t, p, gt = (torch.randn(B, D, generator=g).to(DEV) for _ in range(3))
    if zero_row:
        p[0] = 0
    aes = {k: v.to(DEV) for k, v in cv.aesthetic_state_dict(D, seed=seed + 1).items()}
    miou = torch.rand(B, generator=g).to(DEV) if with_layout else None
    lay = torch.rand(B, generator=g).to(DEV) if with_layout else None
    r, c, a = reward_head(t, p, gt, aes, miou, lay)
    z = torch.zeros(B, device=DEV)
    with torch.no_grad(), true_fp32():
        rr, cc, aa = cv.reward_forward(t, p, gt, aes, z if miou is None else miou, z if lay is None else lay)
    return max(float((r - rr).abs().max()), float((c - cc).abs().max()), float((a - aa).abs().max()))


class _Processor:
    """CLIPProcessor stand-in (host side): uint8 HWC images -> resized / normalised pixel_values."""

    def __init__(self, size):
        self.size = size

    def __call__(self, images=None, return_tensors="pt", **kw):
        x = torch.stack([torch.as_tensor(im).permute(2, 0, 1).float() / 255.0 for im in images])
        x = torch.nn.functional.interpolate(x, size=(self.size, self.size), mode="bilinear", align_corners=False)
        mean = torch.tensor([0.4815, 0.4578, 0.4082]).view(1, 3, 1, 1)
        std = torch.tensor([0.2686, 0.2613, 0.2758]).view(1, 3, 1, 1)
        return dict(pixel_values=(x - mean) / std)


def check_reward_model(which, B=3, seed=0, reference_metrics=True):
    """Reward.forward (captions, generated images, ground-truth images, layouts) against the oracle composition of
    models/policy.py:106-139; the layout terms from the reference's own tools/metrics.py (staged under oracle/_ref)."""
    from ltt_test_stubs import HashTokenizer
    from layoutllm_t2i_b200.reward import COCO_LABELS, Reward
    tcfg = dict(co.tiny_clip_text_config() if which == "tiny" else co.default_clip_text_config())
    vcfg = _vcfg(which)
    tcfg["projection_dim"] = vcfg["projection_dim"]
    sd = dict(co.random_state_dict(tcfg, seed=seed))
    sd.update(cv.random_state_dict(vcfg, seed=seed + 1))
    aes = cv.aesthetic_state_dict(vcfg["projection_dim"], seed=seed + 2)
    metrics = None
    if reference_metrics:
        from oracle import ref_loader as rl
        assert rl.available(), "oracle/_ref is not staged"
        with rl.reference_tree():
            sys.path.insert(0, rl.REF_ROOT)
            try:
                from tools.metrics import compute_docsim, compute_maximum_iou
            finally:
                sys.path.remove(rl.REF_ROOT)
        metrics = (compute_maximum_iou, compute_docsim)
    tok, proc = HashTokenizer(tcfg["vocab_size"]), _Processor(vcfg["image_size"])
    rm = Reward(sd, aes, tok, proc, 0, text_config=tcfg, vision_config=vcfg, labels=COCO_LABELS[:12], metrics=metrics)
    g = torch.Generator().manual_seed(seed + 5)
    captions = ["a person riding a bicycle", "two cars and a bus on the street", "a boat"][:B]
    imgs_pred = [(torch.rand(64, 64, 3, generator=g) * 255).to(torch.uint8).numpy() for _ in range(B)]
    imgs_gt = [(torch.rand(80, 72, 3, generator=g) * 255).to(torch.uint8).numpy() for _ in range(B)]
    box = lambda: sorted(torch.rand(2, generator=g).tolist()) + sorted(torch.rand(2, generator=g).tolist())  # noqa: E731
    lay_gt = [([box(), box()], ["person", "bicycle"]), ([box(), box(), box()], ["car", "car", "bus"]), ([box()], ["boat"])][:B]
    lay_pred = [([box(), box()], ["person", "bike rider"]), ([box(), box()], ["car", "bus"]), ([box()], ["boat"])][:B]   # one open-set label
    reward, clip_r, aes_r, miou, laysim = rm.forward(captions, imgs_pred, imgs_gt, lay_pred, lay_gt, return_parts=True)
    # the attributes the callers read off the reward model (train_rl.py:91,326; data.py:41-54)
    assert rm.to(DEV) is rm and rm.processor is proc and rm.tokenizer is tok
    inputs = tok(captions, padding=True, return_tensors="pt")
    feats = rm.model.get_text_features(**{k: v.to(rm.device) for k, v in inputs.items()})
    assert torch.equal(feats, rm.get_text_features(captions)) and rm.model.projection_dim == vcfg["projection_dim"]
    assert rm.model.get_image_features(pixel_values=proc(images=imgs_pred)["pixel_values"]).shape == (B, vcfg["projection_dim"])
    # ---- oracle composition
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    aesd = {k: v.to(DEV) for k, v in aes.items()}
    with torch.no_grad(), true_fp32():
        t = co.text_features(sdd, tcfg, tok(captions, padding=True)["input_ids"].to(DEV), tok(captions, padding=True)["attention_mask"].to(DEV))
        p = cv.image_features(sdd, vcfg, proc(images=imgs_pred)["pixel_values"].to(DEV))
        q = cv.image_features(sdd, vcfg, proc(images=imgs_gt)["pixel_values"].to(DEV))
        want, want_clip, want_aes = cv.reward_forward(t, p, q, aesd, miou.to(DEV), laysim.to(DEV))
    assert reward.shape == (B,) and torch.isfinite(reward).all()
    return max(float((clip_r - want_clip).abs().max()), float((aes_r - want_aes).abs().max()) * 0.1, float((reward - want).abs().max()))


def check_preprocess(shapes, B=2, seed=0, against="oracle"):
    """Device-side CLIPImageProcessor (resize shortest edge -> 224 with Pillow's BICUBIC, centre crop, rescale, normalise) on
    uint8 HWC device images: fraction of output values that differ from the oracle / from transformers' PIL-backed processor
    (must be 0: integer arithmetic + one float32 table)."""
    import numpy as np
    from oracle import clip_preprocess_oracle as cp
    cfg, sd, enc = vision("tiny224", 0)
    rng = np.random.default_rng(seed)
    bad = 0.0
    for h, w in shapes:
        imgs = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
        got = enc.preprocess(torch.from_numpy(imgs).to(DEV)).cpu().numpy()
        if against == "oracle":
            want = cp.preprocess(list(imgs))
        else:
            from PIL import Image
            from transformers.models.clip import CLIPImageProcessorPil
            want = CLIPImageProcessorPil()(images=[Image.fromarray(i) for i in imgs], return_tensors="np")["pixel_values"]
        assert got.shape == want.shape == (B, 3, 224, 224)
        bad = max(bad, float((got != want).mean()))
    return bad


def check_reward_device_images(seed=0):
    """Reward.forward fed the generated images as a uint8 device tensor (decoder output) == fed through a host processor that
    implements the same pipeline (the oracle's): identical pixel_values -> identical reward bits."""
    import numpy as np
    from ltt_test_stubs import HashTokenizer
    from layoutllm_t2i_b200.reward import COCO_LABELS, Reward
    from oracle import clip_preprocess_oracle as cp
    tcfg = dict(co.tiny_clip_text_config())
    vcfg = dict(cv.tiny_clip_vision_config(), image_size=224)
    tcfg["projection_dim"] = vcfg["projection_dim"]
    sd = dict(co.random_state_dict(tcfg, seed=seed))
    sd.update(cv.random_state_dict(vcfg, seed=seed + 1))
    aes = cv.aesthetic_state_dict(vcfg["projection_dim"], seed=seed + 2)

    class Proc:
        def __call__(self, images=None, return_tensors="pt", **kw):
            return dict(pixel_values=torch.from_numpy(cp.preprocess([np.asarray(i) for i in images])))
    rm = Reward(sd, aes, HashTokenizer(tcfg["vocab_size"]), Proc(), 0, text_config=tcfg, vision_config=vcfg, labels=COCO_LABELS[:4],
                metrics=(lambda a, b: np.full(len(a), 0.25, np.float32), lambda a, b: np.full(len(a), 0.5, np.float32)))
    rng = np.random.default_rng(seed)
    pred = rng.integers(0, 256, (2, 512, 512, 3), dtype=np.uint8)
    gt = rng.integers(0, 256, (2, 300, 260, 3), dtype=np.uint8)
    caps = ["a person on a bicycle", "a car"]
    lay = [([[0.1, 0.1, 0.5, 0.5]], ["person"]), ([[0.2, 0.2, 0.9, 0.9]], ["car"])]
    r_dev = rm.forward(caps, torch.from_numpy(pred).to(DEV), torch.from_numpy(gt).to(DEV), lay, lay)
    r_host = rm.forward(caps, list(pred), list(gt), lay, lay)
    return 0.0 if torch.equal(r_dev, r_host) else float((r_dev - r_host).abs().max()) + 1e-9


def check_reward_vs_reference_class(which="tiny224", B=3, seed=0):
    """The mirror against the reference's OWN `Reward.forward` (models/policy.py:105-139, byte-identical copy under
    oracle/_ref) running eager fp32 on the same GPU: transformers CLIPModel with the same seeded weights, the reference's
    AestheticMLP, transformers' PIL-backed image processor, the reference's tools/metrics.  The class is constructed without
    `from_pretrained` (no network); the only adaptation is version skew -- transformers 5 returns an output object from
    get_*_features where the 4.x API the reference was written for returned the tensor."""
    import numpy as np
    from PIL import Image
    from transformers import CLIPConfig, CLIPModel
    from transformers.models.clip import CLIPImageProcessorPil
    from ltt_test_stubs import HashTokenizer
    from layoutllm_t2i_b200.reward import Reward
    from oracle import ref_loader as rl
    assert rl.available(), "oracle/_ref is not staged"
    with rl.reference_tree():
        sys.path.insert(0, rl.REF_ROOT)
        try:
            from models.policy import Reward as RefReward
            from tools.aesthetic import AestheticMLP
        finally:
            sys.path.remove(rl.REF_ROOT)
    tcfg = dict(co.tiny_clip_text_config() if which != "full" else co.default_clip_text_config())
    vcfg = _vcfg(which)
    tcfg["projection_dim"] = vcfg["projection_dim"]
    sd = dict(co.random_state_dict(tcfg, seed=seed))
    sd.update(cv.random_state_dict(vcfg, seed=seed + 1))
    aes = cv.aesthetic_state_dict(vcfg["projection_dim"], seed=seed + 2)
    tk = ("vocab_size", "max_position_embeddings", "hidden_size", "num_attention_heads", "num_hidden_layers", "intermediate_size",
          "layer_norm_eps", "hidden_act", "eos_token_id")
    vk = ("image_size", "patch_size", "hidden_size", "num_attention_heads", "num_hidden_layers", "intermediate_size", "layer_norm_eps",
          "hidden_act")
    hf = CLIPModel(CLIPConfig(text_config=dict({k: tcfg[k] for k in tk}, bos_token_id=tcfg["vocab_size"] - 2, pad_token_id=1),
                              vision_config={k: vcfg[k] for k in vk}, projection_dim=vcfg["projection_dim"])).to(DEV).eval()
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in k or k == "logit_scale" for k in missing), (missing, unexpected)

    class Hf4:          # transformers 4.x surface of CLIPModel over the installed 5.x
        projection_dim = vcfg["projection_dim"]

        @staticmethod
        def _t(o):
            return o if torch.is_tensor(o) else o.pooler_output

        def get_text_features(self, **kw):
            return self._t(hf.get_text_features(**kw))

        def get_image_features(self, **kw):
            return self._t(hf.get_image_features(**kw))

    tok = HashTokenizer(tcfg["vocab_size"])

    class Enc(dict):
        def to(self, device):
            return Enc({k: v.to(device) for k, v in self.items()})

    class Tok:
        def __call__(self, texts, padding=True, return_tensors="pt"):
            return Enc(tok(texts, padding=padding))
    proc = CLIPImageProcessorPil(size={"shortest_edge": vcfg["image_size"]}, crop_size={"height": vcfg["image_size"], "width": vcfg["image_size"]})
    ref = object.__new__(RefReward)
    torch.nn.Module.__init__(ref)
    ref.tokenizer, ref.processor, ref.model, ref.device = Tok(), proc, Hf4(), DEV
    ref.args = type("A", (), dict(img_dir="x/train2014"))()
    ref.aesthetic_model = AestheticMLP(vcfg["projection_dim"]).to(DEV).eval()
    ref.aesthetic_model.load_state_dict(aes)
    g = torch.Generator().manual_seed(seed + 5)
    captions = ["a person riding a bicycle", "two cars and a bus on the street", "a boat"][:B]
    pred = [(torch.rand(96, 96, 3, generator=g) * 255).to(torch.uint8).numpy() for _ in range(B)]
    gt = [(torch.rand(120, 100, 3, generator=g) * 255).to(torch.uint8).numpy() for _ in range(B)]
    box = lambda: sorted(torch.rand(2, generator=g).tolist()) + sorted(torch.rand(2, generator=g).tolist())  # noqa: E731
    lay_gt = [([box(), box()], ["person", "bicycle"]), ([box(), box(), box()], ["car", "car", "bus"]), ([box()], ["boat"])][:B]
    lay_pred = [([box(), box()], ["person", "bike rider"]), ([box(), box()], ["car", "bus"]), ([box()], ["boat"])][:B]
    with torch.no_grad(), true_fp32():
        ref.emb_labels()
        want = ref.forward(captions, [Image.fromarray(i) for i in pred], [Image.fromarray(i) for i in gt], lay_pred, lay_gt)
    rm = Reward(sd, aes, Tok(), proc, 0, text_config=tcfg, vision_config=vcfg, metrics=None)
    sys.path.insert(0, rl.REF_ROOT)          # Reward's default layout terms: `from tools.metrics import ...` of the caller's checkout
    try:
        with rl.reference_tree():
            got = rm.forward(captions, torch.from_numpy(np.stack(pred)).to(DEV), [Image.fromarray(i) for i in gt], lay_pred, lay_gt)
    finally:
        sys.path.remove(rl.REF_ROOT)
    assert rm.labels == ref.labels
    return float((got - want.to(got.dtype)).abs().max())


# towers: fp16 operands / fp32 accumulate against an fp32 reference over 24 layers (measured values: profiles/r02_reward_parity.txt)
TOL = 3e-3
ALL = [
    ("vision tower tiny vs transformers fixture", check_vision_golden, {}, TOL),
    ("vision tower tiny hidden B=3", check_vision, dict(which="tiny", B=3), TOL),
    ("vision tower tiny embeds B=1", check_vision, dict(which="tiny", B=1, what="embeds"), TOL),
    ("vision tower batch invariance", check_vision_batch_invariance, {}, TOL),
    ("vision tower ViT-L/14 hidden B=2", check_vision, dict(which="full", B=2), TOL),
    ("vision tower ViT-L/14 pooled B=5", check_vision, dict(which="full", B=5, what="pooled"), TOL),
    ("vision tower ViT-L/14 image_embeds B=16", check_vision, dict(which="full", B=16, what="embeds"), TOL),
    ("preprocess 512x512 / 224x224 / 64x80 / 300x260 / 97x211 vs oracle (bit-exact)", check_preprocess,
     dict(shapes=[(512, 512), (224, 224), (64, 80), (300, 260), (97, 211)]), 1e-12),
    ("preprocess 512x512 / 1024x768 vs transformers' PIL-backed processor (bit-exact)", check_preprocess,
     dict(shapes=[(512, 512), (1024, 768)], B=3, seed=1, against="transformers"), 1e-12),
    ("Reward.forward on device uint8 images == host-processor path", check_reward_device_images, {}, 1e-12),
    ("Reward.forward vs the reference's own Reward class (eager fp32 transformers), tiny towers", check_reward_vs_reference_class, {}, 5e-3),
    ("Reward.forward vs the reference's own Reward class, ViT-L/14 towers", check_reward_vs_reference_class, dict(which="full", seed=3), 5e-3),
    ("reward head D=768 B=8", check_reward_head, dict(B=8, D=768), 1e-5),
    ("reward head D=64 B=1, no layout terms", check_reward_head, dict(B=1, D=64, with_layout=False), 1e-5),
    ("reward head zero feature row", check_reward_head, dict(B=3, D=768, zero_row=True), 1e-5),
    # reward = cos + cos + 0.1 aes + ...: absolute error of the scalar (|reward| ~ 1 .. 20)
    ("Reward.forward tiny towers", check_reward_model, dict(which="tiny"), 5e-3),
    ("Reward.forward ViT-L/14 towers", check_reward_model, dict(which="full"), 5e-3),
]


def main():
    for name, fn, kw, tol in ALL:
        try:
            err = fn(**kw)
            torch.cuda.synchronize()
        except Exception as ex:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print(f"EXC  {name:64s} {ex!r}"[:300], flush=True)
            continue
        print(f"{'ok ' if err < tol else 'FAIL'} {name:64s} {err:.3e}  (gate {tol:g})", flush=True)


if __name__ == "__main__":
    main()
