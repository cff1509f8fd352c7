"""VAE decode + image post-processing parity: the sm_100a decoder through the C-ABI
(layoutllm_t2i_b200.vae.VaeDecoder / the drop-in AutoencoderKL) against the UNMODIFIED reference AutoencoderKL
(oracle/_ref, GLIGEN/ldm/models/autoencoder.py + modules/diffusionmodules/model.py) on the same seeded weights and
latents, fp32 and under torch.autocast('cuda', fp16); uint8 images against the callers' expression (txt2img.py:320-323).
Used by tests/test_vae_gpu.py."""
from __future__ import annotations

import numpy as np
import torch

from oracle import ref_loader as rl

DEV = "cuda"
SD_VAE = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
              num_res_blocks=2, attn_resolutions=[], dropout=0.0)
SMALL_VAE = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=64, ch_mult=[1, 2],
                 num_res_blocks=1, attn_resolutions=[], dropout=0.0)
_CACHE = {}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def reference_vae(dd, seed):
    """The reference AutoencoderKL with its own default initialisation under a fixed seed (norm gains / biases perturbed
    so that no affine parameter is trivially 1 / 0)."""
    key = (repr(sorted(dd.items())), seed)
    if key not in _CACHE:
        import contextlib
        import io
        torch.manual_seed(seed)
        with rl.reference_tree(), contextlib.redirect_stdout(io.StringIO()):
            from ldm.models.autoencoder import AutoencoderKL
            m = AutoencoderKL(dd, 4, 0.18215).eval()
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "norm" in n:
                    p.add_(0.1 * torch.randn(p.shape, generator=g))
                elif n.endswith("bias"):
                    p.add_(0.02 * torch.randn(p.shape, generator=g))
        _CACHE[key] = m.to(DEV)
    return _CACHE[key]


def ours(dd, seed, via_dropin):
    key = ("ours", repr(sorted(dd.items())), seed, via_dropin)
    if key not in _CACHE:
        sd = reference_vae(dd, seed).state_dict()
        if via_dropin:
            import os
            import sys
            d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "layoutllm_t2i_b200", "dropin")
            if d not in sys.path:
                sys.path.insert(0, d)
            from ldm.util import instantiate_from_config
            m = instantiate_from_config(dict(target="ldm.models.autoencoder.AutoencoderKL",
                                             params=dict(ddconfig=dd, embed_dim=4, scale_factor=0.18215)))
            m.load_state_dict(sd)            # strict, as txt2img.py:107
            _CACHE[key] = m.to(DEV).eval()
        else:
            from layoutllm_t2i_b200.vae import VaeDecoder
            e = VaeDecoder(dict(ch=dd["ch"], out_ch=dd["out_ch"], ch_mult=dd["ch_mult"], num_res_blocks=dd["num_res_blocks"],
                                z_channels=4, embed_dim=4, scale_factor=0.18215), 0)
            e.load_state_dict(sd)
            e.finalize()
            _CACHE[key] = e
    return _CACHE[key]


def latents(B, h, w, seed=9):
    # sampler outputs are ~ unit-variance latents; the decoder divides by scale_factor itself
    return (0.9 * torch.randn(B, 4, h, w, generator=torch.Generator().manual_seed(seed))).to(DEV)


@torch.no_grad()
def ref_decode(dd, seed, z, autocast):
    m = reference_vae(dd, seed)
    if autocast:
        with torch.autocast("cuda", dtype=torch.float16):
            return m.decode(z).float()
    with rl.true_fp32():
        return m.decode(z).float()


def check_decode(dd, seed, B, h, w, autocast=True, via_dropin=False):
    z = latents(B, h, w)
    r = ref_decode(dd, seed, z, autocast)
    o = ours(dd, seed, via_dropin).decode(z)
    assert o.shape == r.shape and torch.isfinite(o).all()
    return rel(o, r)


def check_uint8(dd, seed, B, h, w):
    """(1) the fused uint8 image equals the callers' expression applied to our own decode output (exact);
    (2) against the reference pipeline it differs by at most one grey level on >= 99.5% of the bytes.  Returns the
    fraction of bytes violating (1) or (2)."""
    z = latents(B, h, w, seed=21)
    dec = ours(dd, seed, False)
    img, u8 = dec.decode(z, images_u8=True)
    host = dec.decode_to_uint8(z)
    assert host.is_pinned() and torch.equal(host, u8.cpu())
    mine = np.stack([(torch.clamp(s, min=-1, max=1) * 0.5 + 0.5).cpu().numpy().transpose(1, 2, 0) * 255 for s in img]).astype(np.uint8)
    exact_bad = float((mine != u8.cpu().numpy()).mean())
    r = ref_decode(dd, seed, z, True)
    theirs = np.stack([(torch.clamp(s, min=-1, max=1) * 0.5 + 0.5).cpu().numpy().transpose(1, 2, 0) * 255 for s in r]).astype(np.uint8)
    off = np.abs(theirs.astype(np.int32) - u8.cpu().numpy().astype(np.int32))
    print(f"uint8 images: exact-vs-own-expression mismatch {exact_bad:.2e}; vs reference pipeline: {float((off > 1).mean()):.2e} of bytes differ by > 1 level, max {int(off.max())}")
    return max(exact_bad, float((off > 1).mean()) - 0.005)


ALL = [
    ("VAE decode small (64ch, 16x16 latent, B=2) vs reference fp32", check_decode, dict(dd=SMALL_VAE, seed=1, B=2, h=16, w=16, autocast=False), 5e-3),
    ("VAE decode small vs reference fp16 autocast", check_decode, dict(dd=SMALL_VAE, seed=1, B=2, h=16, w=16), 5e-3),
    ("VAE decode small, 24x16 latent, B=3", check_decode, dict(dd=SMALL_VAE, seed=1, B=3, h=24, w=16), 5e-3),
    ("VAE decode SD config 64x64 -> 512x512, B=1 vs reference fp16 autocast", check_decode, dict(dd=SD_VAE, seed=2, B=1, h=64, w=64), 5e-3),
    ("VAE decode SD config 64x64, B=1 vs reference fp32", check_decode, dict(dd=SD_VAE, seed=2, B=1, h=64, w=64, autocast=False), 5e-3),
    ("drop-in AutoencoderKL.decode (strict load_state_dict), SD config B=2", check_decode, dict(dd=SD_VAE, seed=2, B=2, h=64, w=64, via_dropin=True), 5e-3),
    ("VAE decode SD config 96x96 -> 768x768, B=1", check_decode, dict(dd=SD_VAE, seed=2, B=1, h=96, w=96), 5e-3),
    ("uint8 HWC images (fused post-processing + one pinned copy), small", check_uint8, dict(dd=SMALL_VAE, seed=1, B=2, h=16, w=16), 1e-9),
    ("uint8 HWC images, SD config 512x512", check_uint8, dict(dd=SD_VAE, seed=2, B=1, h=64, w=64), 1e-9),
]
