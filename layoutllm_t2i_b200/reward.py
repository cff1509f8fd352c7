"""The reward of the RL loop on the sm_100a library: mirror of ``Reward`` in the reference's
``models/policy.py:36-139`` -- CLIP text / image features, cosine rewards, aesthetic predictor, layout terms.

What runs where: the captions go through the text tower (``ltt_clip_encode`` -> ``text_embeds``), the generated AND the
ground-truth images through ONE pass of the vision tower (``ltt_clip_vision_encode`` on the 2B stacked images -- the
reference runs two eager fp32 ``get_image_features`` calls), and a single kernel (``ltt_reward_head``) evaluates the
normalisations, both cosines, the five-layer aesthetic predictor and the final weighted sum.  Host side, as in the
reference: the tokenizer / image processor (third-party), and the two layout terms -- ``compute_maximum_iou`` /
``compute_docsim`` are numpy / scipy assignment problems over a handful of boxes (the reference's ``tools/metrics.py``, imported
from the caller's checkout or passed in).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from .clip import ClipModelAdapter, ClipTextEncoder, ClipVisionEncoder

COCO_LABELS = ['person', 'bicycle', 'car', 'motorcycle', 'airplane', 'bus', 'train', 'truck', 'boat', 'traffic light', 'fire hydrant',
               'stop sign', 'parking meter', 'bench', 'bird', 'cat', 'dog', 'horse', 'sheep', 'cow', 'elephant', 'bear', 'zebra',
               'giraffe', 'backpack', 'umbrella', 'handbag', 'tie', 'suitcase', 'frisbee', 'skis', 'snowboard', 'sports ball', 'kite',
               'baseball bat', 'baseball glove', 'skateboard', 'surfboard', 'tennis racket', 'bottle', 'wine glass', 'cup', 'fork',
               'knife', 'spoon', 'bowl', 'banana', 'apple', 'sandwich', 'orange', 'broccoli', 'carrot', 'hot dog', 'pizza', 'donut',
               'cake', 'chair', 'couch', 'potted plant', 'bed', 'dining table', 'toilet', 'tv', 'laptop', 'mouse', 'remote',
               'keyboard', 'cell phone', 'microwave', 'oven', 'toaster', 'sink', 'refrigerator', 'book', 'clock', 'vase', 'scissors',
               'teddy bear', 'hair drier', 'toothbrush']          # the closed label set of Reward.emb_labels (policy.py:56-66)
_AES_LAYERS = (0, 2, 4, 6, 7)                                      # the nn.Linear slots of AestheticMLP.layers (aesthetic.py:15-27)


def reward_head(txt: torch.Tensor, pred: torch.Tensor, gt: torch.Tensor, aes_sd: Dict[str, torch.Tensor],
                miou: Optional[torch.Tensor] = None, laysim: Optional[torch.Tensor] = None):
    """(reward, clip_reward, aes_reward), each [B], from the three [B, D] feature matrices (policy.py:115-139)."""
    dev = pred.device
    B, D = pred.shape
    t, p, g = (x.detach().to(device=dev, dtype=torch.float32).contiguous() for x in (txt, pred, gt))
    ws = [aes_sd[f"layers.{i}.weight"].detach().to(device=dev, dtype=torch.float32).contiguous() for i in _AES_LAYERS]
    bs = [aes_sd[f"layers.{i}.bias"].detach().to(device=dev, dtype=torch.float32).contiguous() for i in _AES_LAYERS]
    dims = (C.c_int * 6)(D, *[int(w.shape[0]) for w in ws])
    for w, i, o in zip(ws, list(dims)[:-1], list(dims)[1:]):
        if tuple(w.shape) != (o, i):
            raise L.LttError(f"reward_head: aesthetic layer of shape {tuple(w.shape)} does not chain ({i} -> {o})")
    wp = (C.c_void_p * 5)(*[w.data_ptr() for w in ws])
    bp = (C.c_void_p * 5)(*[b.data_ptr() for b in bs])
    mi = None if miou is None else miou.detach().to(device=dev, dtype=torch.float32).contiguous()
    ls = None if laysim is None else laysim.detach().to(device=dev, dtype=torch.float32).contiguous()
    out = torch.empty(3, B, device=dev)
    with torch.cuda.device(dev):
        L.check(L.lib().ltt_reward_head(L.ptr(t), L.ptr(p), L.ptr(g), B, D, wp, bp, dims, L.ptr(mi), L.ptr(ls), L.ptr(out[0]),
                                        L.ptr(out[1]), L.ptr(out[2]), L.stream_ptr()), "ltt_reward_head")
    return out[0], out[1], out[2]


class Reward:
    """Reward model of train_rl.py (``Reward(args.model_config, args.aesthetic_ckpt, args, device)``, policy.py:36-50).

    Built from state dicts instead of ``from_pretrained`` names: ``clip_sd`` = ``CLIPModel.state_dict()`` (both towers and
    both projections), ``aesthetic_sd`` = the AestheticMLP checkpoint; ``tokenizer`` / ``processor`` are the caller's
    ``AutoTokenizer`` / ``AutoProcessor`` (host side).  ``metrics`` = (compute_maximum_iou, compute_docsim); by default they
    are imported from the reference checkout's ``tools.metrics``."""

    def __init__(self, clip_sd: Dict[str, torch.Tensor], aesthetic_sd: Dict[str, torch.Tensor], tokenizer, processor, device=0,
                 text_config: Optional[dict] = None, vision_config: Optional[dict] = None, labels: Sequence[str] = COCO_LABELS,
                 metrics: Optional[Sequence[Callable]] = None):
        pd = int(clip_sd["text_projection.weight"].shape[0])
        self.text = ClipTextEncoder(dict(text_config or {}, projection_dim=pd), device)
        self.vision = ClipVisionEncoder(dict(vision_config or {}, projection_dim=pd), device)
        self.text.load_state_dict(clip_sd)
        self.vision.load_state_dict(clip_sd)
        self.text.finalize()
        self.vision.finalize()
        self.device = self.text.device
        self.projection_dim = pd
        self.aesthetic_sd = {k: v.detach().to(self.device, torch.float32) for k, v in aesthetic_sd.items()}
        self.tokenizer, self.processor = tokenizer, processor
        self._metrics = tuple(metrics) if metrics is not None else None
        # what the callers read off the reward model: `.model` / `.processor` go to run_batch_images as clip_model /
        # clip_processor (train_rl.py:91), `.tokenizer` / `.model.get_text_features` / `.device` to load_data (data.py:41-54)
        self.model = ClipModelAdapter(self.text, self.vision)
        self.labels = list(labels)
        self.label2index = {l: i for i, l in enumerate(self.labels)}
        self.emb_labels()

    def to(self, *a, **k):          # train_rl.py:326 (`reward_model.to(device)`): the towers already live on `device`
        return self

    def eval(self):
        return self

    # ---- policy.py:53-75
    def get_text_features(self, texts: Sequence[str]) -> torch.Tensor:
        ids = self.tokenizer(list(texts), padding=True, return_tensors="pt")["input_ids"]
        return self.text.encode_ids(ids, want_hidden=False, want_embeds=True)[2]

    def get_image_features(self, images) -> torch.Tensor:
        return self.vision.encode(self.processor(images=images, return_tensors="pt")["pixel_values"])[2]

    def _pixel_values(self, images) -> torch.Tensor:
        """uint8 [B, H, W, 3] tensors already on the device (the decoder's images) are preprocessed there -- the RL loop's
        generated images never visit the host; anything else (PIL images, arrays) goes through the caller's processor."""
        if torch.is_tensor(images) and images.dtype == torch.uint8 and images.is_cuda:
            return self.vision.preprocess(images)
        return self.processor(images=images, return_tensors="pt")["pixel_values"].to(self.device)

    def emb_labels(self) -> None:
        self.labels_emb = torch.nn.functional.normalize(self.get_text_features(self.labels), dim=-1)

    # ---- policy.py:77-102 (host bookkeeping, unchanged in substance)
    def label_to_id(self, layouts):
        return [(np.array(boxes), np.array([self.label2index[l] for l in labels])) for boxes, labels in layouts]

    def nn_close_set(self, layouts):
        out = []
        for boxes, labels in layouts:
            new = []
            for label in labels:
                if label in self.label2index:
                    new.append(label)
                else:      # nearest closed-set label in CLIP text space
                    emb = torch.nn.functional.normalize(self.get_text_features([label]), dim=-1)
                    new.append(self.labels[int((emb @ self.labels_emb.t()).flatten().argmax())])
            out.append((boxes, new))
        return out

    def _layout_terms(self, layout_pred, layout_gt):
        if self._metrics is None:
            from tools.metrics import compute_docsim, compute_maximum_iou      # the reference checkout's own host code
            self._metrics = (compute_maximum_iou, compute_docsim)
        pred_id = self.label_to_id(self.nn_close_set(layout_pred))
        gt_id = self.label_to_id(layout_gt)
        miou = torch.from_numpy(np.asarray(self._metrics[0](gt_id, pred_id), dtype=np.float32))
        laysim = torch.from_numpy(np.asarray(self._metrics[1](gt_id, pred_id), dtype=np.float32))
        return miou, laysim

    # ---- policy.py:105-139
    @torch.no_grad()
    def forward(self, captions: List[str], imgs_pred, imgs_gt, layout_pred, layout_gt, return_parts: bool = False):
        txt = self.get_text_features(captions)
        px = torch.cat([self._pixel_values(imgs_pred), self._pixel_values(imgs_gt)])
        emb = self.vision.encode(px)[2]                                   # ONE vision pass over generated + ground-truth images
        B = txt.shape[0]
        miou, laysim = self._layout_terms(layout_pred, layout_gt)
        reward, clip_r, aes_r = reward_head(txt, emb[:B], emb[B:], self.aesthetic_sd, miou, laysim)
        return (reward, clip_r, aes_r, miou, laysim) if return_parts else reward

    __call__ = forward
