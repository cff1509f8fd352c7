"""Conditioning prep on the sm_100a library (``ltt_clip`` of include/ltt_b200.h): the CLIP text tower and the host-side
assembly of the sampler's conditioning tensors in ONE batched pass.

Mirrors, in the reference (/root/reference): ``FrozenCLIPEmbedder`` GLIGEN/ldm/modules/encoders/modules.py:144-182,
``get_clip_feature`` txt2img.py:126-156, ``prepare_batch`` :173-209, ``prepare_relation_phrases`` :213-244 and the
conditioning block of ``generate_one_image`` :268-277.  The reference encodes every string with its own eager fp32 call
(prompt, "", each box phrase -- with a dummy 224 x 224 vision pass each --, the relation phrases); here all strings of an
image are rows of one [B, 77] id batch.  That is exact, not an approximation: attention is causal, so the rows up to the
end-of-text token -- and the pooled vector read there -- do not depend on the padding behind it
(tests/test_clip_oracle_cpu.py::test_batched_padded_pass_equals_per_phrase_calls).

No arithmetic here: tokenisation stays with the caller's tokenizer (third-party, host side), PyTorch-owned device
pointers are forwarded to the C-ABI.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import _lib as L


def default_clip_text_config() -> dict:
    """Text tower of openai/clip-vit-large-patch14 (FrozenCLIPEmbedder's ``version``, modules.py:146)."""
    return dict(vocab_size=49408, max_position_embeddings=77, hidden_size=768, num_attention_heads=12, num_hidden_layers=12,
                intermediate_size=3072, layer_norm_eps=1e-5, hidden_act="quick_gelu", projection_dim=0, eos_token_id=2)


_PREFIXES = ("transformer.text_model.", "text_model.")
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)        # transformers OPENAI_CLIP_MEAN / OPENAI_CLIP_STD
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class ClipTextEncoder:
    """One text tower per CUDA device; stream ordered on the current torch stream."""

    def __init__(self, cfg: Optional[dict] = None, device=0):
        if not torch.cuda.is_available():
            raise L.LttError("layoutllm_t2i_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        base = default_clip_text_config()
        self.cfg = dict(base, **{k: v for k, v in (cfg or {}).items() if k in base})
        if self.cfg["hidden_act"] != "quick_gelu":
            raise L.LttError(f"ClipTextEncoder: hidden_act {self.cfg['hidden_act']!r} is not supported (quick_gelu only)")
        c = L.ClipConfig()
        c.vocab, c.max_pos, c.hidden = self.cfg["vocab_size"], self.cfg["max_position_embeddings"], self.cfg["hidden_size"]
        c.heads, c.layers, c.ffn = self.cfg["num_attention_heads"], self.cfg["num_hidden_layers"], self.cfg["intermediate_size"]
        c.eps, c.act, c.proj_dim = float(self.cfg["layer_norm_eps"]), 0, int(self.cfg["projection_dim"] or 0)
        c.eos_token_id = int(self.cfg["eos_token_id"])
        self._h = C.c_void_p()
        self._lib = L.lib()
        L.check(self._lib.ltt_clip_create(C.byref(c), self.device.index, C.byref(self._h)), "ltt_clip_create")
        self._finalized = False

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Accepts the state_dict of FrozenCLIPEmbedder ("transformer.text_model. ..."), of transformers CLIPTextModel
        ("text_model. ...") or of CLIPModel (text tower + "text_projection.weight"; vision keys are skipped)."""
        for k, v in sd.items():
            key = None
            for p in _PREFIXES:
                if k.startswith(p):
                    key = "text_model." + k[len(p):]
                    break
            if k == "text_projection.weight" and self.cfg["projection_dim"]:
                key = k
            if key is None or key.endswith("position_ids"):
                continue
            t = v.detach().to(torch.float32).contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            L.check(self._lib.ltt_clip_load_param(self._h, key.encode(), L.ptr(t), shape, t.dim(), 0 if t.is_cuda else 1),
                    f"ltt_clip_load_param({key})")
        self._finalized = False

    def finalize(self) -> None:
        L.check(self._lib.ltt_clip_finalize(self._h), "ltt_clip_finalize")
        self._finalized = True

    def encode_ids(self, ids: torch.Tensor, want_hidden: bool = True, want_embeds: bool = False):
        """ids [B, L] (L <= 77) -> (last_hidden_state [B, L, W] or None, pooler_output [B, W][, text_embeds [B, P]])."""
        if ids.dim() != 2 or ids.shape[1] < 1 or ids.shape[1] > self.cfg["max_position_embeddings"]:
            raise L.LttError(f"ClipTextEncoder: ids must be [B, L <= {self.cfg['max_position_embeddings']}], got {tuple(ids.shape)}")
        # host ids (what a tokenizer returns) are range-checked here; ids already on the device are clamped by the kernel
        # instead -- checking them would cost two reductions and a host synchronisation per call
        if not ids.is_cuda and ids.numel() and (int(ids.min()) < 0 or int(ids.max()) >= self.cfg["vocab_size"]):
            raise L.LttError("ClipTextEncoder: token id outside the vocabulary")
        if not self._finalized:
            self.finalize()
        ii = ids.detach().to(device=self.device, dtype=torch.int32).contiguous()
        B, Lq = ii.shape
        W = self.cfg["hidden_size"]
        hid = torch.empty(B, Lq, W, device=self.device) if want_hidden else None
        pooled = torch.empty(B, W, device=self.device)
        emb = torch.empty(B, self.cfg["projection_dim"], device=self.device) if want_embeds else None
        with torch.cuda.device(self.device):
            L.check(self._lib.ltt_clip_encode(self._h, L.ptr(ii), B, Lq, L.ptr(hid), L.ptr(pooled), L.ptr(emb), L.stream_ptr()),
                    "ltt_clip_encode")
        self._keep = ii
        return (hid, pooled, emb) if want_embeds else (hid, pooled)

    @property
    def launch_count(self) -> int:
        return int(self._lib.ltt_clip_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ltt_clip_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def default_clip_vision_config() -> dict:
    """Vision tower of openai/clip-vit-large-patch14 (the ``model_config`` of the reward model, train_rl.py:325)."""
    return dict(image_size=224, patch_size=14, hidden_size=1024, num_attention_heads=16, num_hidden_layers=24,
                intermediate_size=4096, layer_norm_eps=1e-5, hidden_act="quick_gelu", projection_dim=768)


class ClipVisionEncoder:
    """CLIP vision tower (``ltt_clip_vision`` of include/ltt_b200.h): transformers ``CLIPVisionTransformer`` +
    ``visual_projection`` as ``Reward.forward`` calls them (/root/reference/models/policy.py:111-114)."""

    def __init__(self, cfg: Optional[dict] = None, device=0):
        if not torch.cuda.is_available():
            raise L.LttError("layoutllm_t2i_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        base = default_clip_vision_config()
        self.cfg = dict(base, **{k: v for k, v in (cfg or {}).items() if k in base})
        if self.cfg["hidden_act"] != "quick_gelu":
            raise L.LttError(f"ClipVisionEncoder: hidden_act {self.cfg['hidden_act']!r} is not supported (quick_gelu only)")
        c = L.ClipVisionConfig()
        c.image_size, c.patch, c.hidden = self.cfg["image_size"], self.cfg["patch_size"], self.cfg["hidden_size"]
        c.heads, c.layers, c.ffn = self.cfg["num_attention_heads"], self.cfg["num_hidden_layers"], self.cfg["intermediate_size"]
        c.eps, c.act, c.proj_dim = float(self.cfg["layer_norm_eps"]), 0, int(self.cfg["projection_dim"] or 0)
        self.tokens = (self.cfg["image_size"] // self.cfg["patch_size"]) ** 2 + 1
        self._h = C.c_void_p()
        self._lib = L.lib()
        L.check(self._lib.ltt_clip_vision_create(C.byref(c), self.device.index, C.byref(self._h)), "ltt_clip_vision_create")
        self._finalized = False

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """Accepts the state_dict of transformers CLIPVisionModel or CLIPModel (text keys are skipped)."""
        for k, v in sd.items():
            if not (k.startswith("vision_model.") or (k == "visual_projection.weight" and self.cfg["projection_dim"])):
                continue
            if k.endswith("position_ids"):
                continue
            t = v.detach().to(torch.float32).contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            L.check(self._lib.ltt_clip_vision_load_param(self._h, k.encode(), L.ptr(t), shape, t.dim(), 0 if t.is_cuda else 1),
                    f"ltt_clip_vision_load_param({k})")
        self._finalized = False

    def finalize(self) -> None:
        L.check(self._lib.ltt_clip_vision_finalize(self._h), "ltt_clip_vision_finalize")
        self._finalized = True

    def encode(self, pixel_values: torch.Tensor, want_hidden: bool = False, want_embeds: bool = True):
        """pixel_values [B, 3, S, S] (CLIPProcessor output) -> (last_hidden_state or None, pooler_output, image_embeds or None)."""
        S = self.cfg["image_size"]
        if pixel_values.dim() != 4 or tuple(pixel_values.shape[1:]) != (3, S, S):
            raise L.LttError(f"ClipVisionEncoder: pixel_values must be [B, 3, {S}, {S}], got {tuple(pixel_values.shape)}")
        if want_embeds and not self.cfg["projection_dim"]:
            raise L.LttError("ClipVisionEncoder: image_embeds requested but the tower was built without visual_projection")
        if not self._finalized:
            self.finalize()
        px = pixel_values.detach().to(device=self.device, dtype=torch.float32).contiguous()
        B, W = px.shape[0], self.cfg["hidden_size"]
        hid = torch.empty(B, self.tokens, W, device=self.device) if want_hidden else None
        pooled = torch.empty(B, W, device=self.device)
        emb = torch.empty(B, self.cfg["projection_dim"], device=self.device) if want_embeds else None
        with torch.cuda.device(self.device):
            L.check(self._lib.ltt_clip_vision_encode(self._h, L.ptr(px), B, L.ptr(hid), L.ptr(pooled), L.ptr(emb), L.stream_ptr()),
                    "ltt_clip_vision_encode")
        self._keep = px
        return hid, pooled, emb

    def preprocess(self, images_u8: torch.Tensor, mean=CLIP_MEAN, std=CLIP_STD) -> torch.Tensor:
        """[B, H, W, 3] uint8 images ON THE DEVICE (what ``VaeDecoder.decode(..., images_u8=True)`` returns) -> pixel_values
        [B, 3, S, S]: the CLIPImageProcessor pipeline of ``Reward.forward`` (models/policy.py:109-112) without the round trip
        through PIL on the host; bit-exact against Pillow's BICUBIC resize + the float32 rescale / normalise."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[-1] != 3:
            raise L.LttError(f"ClipVisionEncoder.preprocess: expected uint8 [B, H, W, 3], got {images_u8.dtype} {tuple(images_u8.shape)}")
        im = images_u8.detach().to(self.device).contiguous()
        B, H, W, _ = im.shape
        S = self.cfg["image_size"]
        out = torch.empty(B, 3, S, S, device=self.device)
        m, s = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
        with torch.cuda.device(self.device):
            L.check(self._lib.ltt_clip_vision_preprocess(self._h, L.ptr(im), B, H, W, m, s, L.ptr(out), L.stream_ptr()),
                    "ltt_clip_vision_preprocess")
        self._keep_img = im
        return out

    @property
    def launch_count(self) -> int:
        return int(self._lib.ltt_clip_vision_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ltt_clip_vision_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class ClipModelAdapter:
    """Stands where the reference's callers pass transformers' ``CLIPModel`` for its TEXT tower: ``model(**inputs)
    .text_model_output.pooler_output`` (get_clip_feature, txt2img.py:147-156; GLIGEN/interface.py has the same function) and
    ``model.get_text_features(**inputs)`` (extract_text_feat, txt2img.py:454-457).  The dummy ``pixel_values`` the callers
    attach are ignored -- no vision pass is run -- and so is ``attention_mask``: under the causal mask the rows up to the
    end-of-text token, which is where the pooled vector is read, do not see the padding.  With a vision tower attached,
    ``get_image_features`` serves the reward model (models/policy.py:111-114); the grounding-by-image branch of
    ``get_clip_feature`` (``is_image=True`` with a real image) is outside the text-to-image path and raises."""

    def __init__(self, encoder: ClipTextEncoder, vision: Optional["ClipVisionEncoder"] = None):
        self.encoder = encoder
        self.vision = vision
        self.projection_dim = int(encoder.cfg["projection_dim"] or 0)        # Reward reads model.projection_dim (policy.py:45)

    @property
    def device(self):
        return self.encoder.device

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def __call__(self, input_ids=None, attention_mask=None, pixel_values=None, **kw):
        if input_ids is None:
            raise L.LttError("ClipModelAdapter: the text tower needs input_ids")
        if tuple(input_ids.shape) == (1, 4) and input_ids.flatten().tolist() == [0, 1, 2, 3]:     # the placeholder ids of the image branch (txt2img.py:139)
            raise L.LttError("ClipModelAdapter: image features (CLIP vision tower) are not part of the text-to-image path")
        want = bool(self.encoder.cfg["projection_dim"])
        out = self.encoder.encode_ids(input_ids, want_embeds=want)
        return SimpleNamespace(text_model_output=SimpleNamespace(last_hidden_state=out[0], pooler_output=out[1]),
                               text_embeds=out[2] if want else None, image_embeds=None)

    def get_text_features(self, input_ids=None, attention_mask=None, **kw):
        if not self.encoder.cfg["projection_dim"]:
            raise L.LttError("ClipModelAdapter.get_text_features: the tower was built without text_projection (projection_dim = 0)")
        return self.encoder.encode_ids(input_ids, want_hidden=False, want_embeds=True)[2]


    def get_image_features(self, pixel_values=None, **kw):
        """CLIPModel.get_image_features (models/policy.py:111-114)."""
        if self.vision is None:
            raise L.LttError("ClipModelAdapter.get_image_features: no vision tower was attached")
        return self.vision.encode(pixel_values)[2]


def relation_phrases(graph: dict, max_relas: int = 5) -> List[str]:
    """The string list ``prepare_relation_phrases`` (txt2img.py:218-238) encodes for a parsed scene graph
    (``sng_parser.parse(prompt)``: the parser is third-party and stays with the caller): "PAD", then every
    "subject relation object" triplet -- listed twice, as the reference's two loops do -- cut to ``max_relas``.
    Empty when the graph has no relations (the reference then returns zeros without encoding anything)."""
    rels = graph.get("relations") or []
    if not rels:
        return []
    ent = graph["entities"]
    trip = [" ".join([ent[r["subject"]]["lemma_head"], r["relation"], ent[r["object"]]["lemma_head"]]) for r in rels]
    return (["PAD"] + trip + trip)[:max_relas]


def prepare_conditioning(encoder: ClipTextEncoder, tokenize: Callable[[Sequence[str]], torch.Tensor], prompt: str,
                         phrases: Sequence[Optional[str]], locations: Sequence[Sequence[float]], relations: Sequence[str] = (),
                         batch: int = 1, max_objs: int = 30, max_relas: int = 5, negative_prompt: str = "",
                         text_mask=None) -> Dict[str, torch.Tensor]:
    """Everything ``generate_one_image`` computes before the sampler (txt2img.py:268-277,296) with ONE text-tower pass.

    ``tokenize(list of str) -> [n, 77] ids`` is FrozenCLIPEmbedder's tokenizer call (modules.py:158-159: truncation,
    max_length 77, padding="max_length").  Returns context / uc [batch, 77, W], relations [batch, max_relas, W] and the
    grounding batch of ``prepare_batch`` (boxes, masks, text_masks, text_embeddings; image entries zero -- the text-to-image
    path passes no reference images).  Requires the phrase tower and the prompt tower to hold the same weights, which is the
    reference's set-up (both are ``openai/clip-vit-large-patch14``)."""
    phrases = list(phrases)
    if len(phrases) > max_objs:
        raise ValueError(f"prepare_conditioning: {len(phrases)} phrases exceed max_objs = {max_objs}")
    n_boxes = min(len(phrases), len(locations))            # prepare_batch zips locations with the phrase features
    relations = list(relations)[:max_relas]
    live = [i for i, p in enumerate(phrases[:n_boxes]) if p is not None]
    strings = [prompt, negative_prompt] + [phrases[i] for i in live] + relations
    hid, pooled = encoder.encode_ids(tokenize(strings))
    W = hid.shape[-1]
    dev = hid.device
    n_ph = len(live)
    # the small bookkeeping tensors are assembled on the host and moved once
    boxes_h = torch.zeros(max_objs, 4)
    masks_h = torch.zeros(max_objs)
    text_masks_h = torch.zeros(max_objs)
    if n_boxes:
        boxes_h[:n_boxes] = torch.as_tensor([list(map(float, b)) for b in locations[:n_boxes]], dtype=torch.float32)
        masks_h[:n_boxes] = 1
    if n_ph:
        text_masks_h[live] = 1
    tm = torch.ones(max_objs)
    if text_mask is not None:                  # complete_mask (txt2img.py:160-170)
        if isinstance(text_mask, (int, float)):
            tm = tm * text_mask
        else:
            for i, v in enumerate(text_mask):
                tm[i] = v
    boxes, masks, text_masks = boxes_h.to(dev), masks_h.to(dev), (text_masks_h * tm).to(dev)
    text_emb = torch.zeros(max_objs, W, device=dev)
    if n_ph:
        text_emb[torch.as_tensor(live, device=dev)] = pooled[2:2 + n_ph]
    rel = torch.zeros(max_relas, W, device=dev)
    if relations:
        rel[:len(relations)] = pooled[2 + n_ph:]
    rep = lambda t: t.unsqueeze(0).repeat(batch, *([1] * t.dim()))  # noqa: E731
    return dict(context=rep(hid[0]), uc=rep(hid[1]), relations=rep(rel),
                boxes=rep(boxes), masks=rep(masks), text_masks=rep(text_masks), image_masks=rep(torch.zeros(max_objs, device=dev)),
                text_embeddings=rep(text_emb), image_embeddings=rep(torch.zeros(max_objs, W, device=dev)))
