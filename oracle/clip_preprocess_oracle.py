"""CPU restatement of the image preprocessing in front of the CLIP vision tower.

TEST INFRASTRUCTURE ONLY (imported by ``tests/`` and timing tools, never by the product package).

Reference call site: ``Reward.forward`` /root/reference/models/policy.py:109-112 -- ``self.processor(images=imgs, return_tensors="pt")``
with ``AutoProcessor.from_pretrained('openai/clip-vit-large-patch14')``, i.e. transformers ``CLIPImageProcessor``: convert to
RGB, resize the shortest edge to 224 with PIL BICUBIC, centre-crop 224 x 224, rescale by 1/255, normalise with the CLIP mean /
std.  The arithmetic lives in two third-party dependencies: Pillow (``Image.resize`` -> ``src/libImaging/Resample.c``: separable
convolution with an anti-aliasing support scaled by the down-scale factor, coefficients normalised in double precision and
quantised to 22 fractional bits, 8-bit rounding after EACH pass, horizontal pass first) and transformers / numpy
(``rescale`` in float64 -> float32, ``normalize`` in float32).  Both are restated here in integer / IEEE arithmetic and pinned
bit-for-bit to the installed Pillow (12.2) and transformers (5.5) by ``tests/test_clip_oracle_cpu.py``
(`test_preprocess_*`) and the committed fixture ``tests/golden/clip_preprocess.npz``.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
PRECISION_BITS = 32 - 8 - 2            # Resample.c


def _bicubic(x: float) -> float:
    a = -0.5                           # Resample.c bicubic_filter
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full-image box -> (bounds [out, 2], kk [out, ksize] int32, ksize)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One resampling pass over `axis` of a uint8 [H, W, C] image (ImagingResampleHorizontal / Vertical_8bpc)."""
    bounds, kk, ksize = precompute_coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = np.tensordot(kk[xx, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0)) + (1 << (PRECISION_BITS - 1))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bicubic(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Image.fromarray(img).resize((out_w, out_h), BICUBIC) for uint8 RGB: horizontal pass, 8-bit rounding, vertical pass."""
    if img.shape[1] != out_w:
        img = _pass(img, out_w, axis=1)
    if img.shape[0] != out_h:
        img = _pass(img, out_h, axis=0)
    return img


def output_geometry(h: int, w: int, size: int = 224) -> Tuple[int, int, int, int]:
    """(resized_h, resized_w, crop_top, crop_left): shortest edge -> size (long edge int(size * long / short)), centre crop."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    rh, rw = (new_long, new_short) if w <= h else (new_short, new_long)
    return rh, rw, (rh - size) // 2, (rw - size) // 2


def normalise_lut(mean=CLIP_MEAN, std=CLIP_STD) -> np.ndarray:
    """[3, 256] float32: rescale (uint8 * (1 / 255) in float64, cast to float32) then (x - mean) / std in float32."""
    v = (np.arange(256, dtype=np.uint8).astype(np.float64) * (1 / 255)).astype(np.float32)
    m, s = np.array(mean, np.float32), np.array(std, np.float32)
    return ((v[None, :] - m[:, None]) / s[:, None]).astype(np.float32)


def preprocess(images, size: int = 224) -> np.ndarray:
    """list of uint8 [H, W, 3] arrays -> pixel_values float32 [B, 3, size, size] (CLIPImageProcessor.preprocess)."""
    lut = normalise_lut()
    out = []
    for im in images:
        im = np.asarray(im)
        rh, rw, top, left = output_geometry(im.shape[0], im.shape[1], size)
        r = pil_resize_bicubic(im, rw, rh)[top:top + size, left:left + size]
        out.append(np.stack([lut[c][r[:, :, c]] for c in range(3)]))
    return np.stack(out)
