"""Python handle on the sm_100a VAE decoder (``ltt_vae`` of include/ltt_b200.h): AutoencoderKL.decode of the reference
(GLIGEN/ldm/models/autoencoder.py:40-44) plus the callers' uint8 / HWC image conversion (txt2img.py:320-323) fused into
the last convolution.  No arithmetic here: PyTorch-owned device pointers are forwarded to the C-ABI."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L


def default_vae_config() -> dict:
    """ddconfig / embed_dim / scale_factor of the LayoutLLM-T2I autoencoder (reference GLIGEN/configs/coco2014.yaml:33-52)."""
    return dict(ch=128, out_ch=3, ch_mult=[1, 2, 4, 4], num_res_blocks=2, z_channels=4, embed_dim=4, scale_factor=0.18215)


class VaeDecoder:
    """One decoder per (AutoencoderKL, CUDA device); stream ordered on the current torch stream."""

    def __init__(self, cfg: dict, device=0):
        if not torch.cuda.is_available():
            raise L.LttError("layoutllm_t2i_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.cfg = dict(default_vae_config(), **{k: v for k, v in cfg.items() if k in default_vae_config()})
        c = L.VaeConfig()
        c.ch, c.out_ch, c.num_res_blocks = self.cfg["ch"], self.cfg["out_ch"], self.cfg["num_res_blocks"]
        mult = list(self.cfg["ch_mult"])
        c.n_levels = len(mult)
        for i, v in enumerate(mult):
            c.ch_mult[i] = int(v)
        c.z_channels, c.embed_dim, c.scale_factor = self.cfg["z_channels"], self.cfg["embed_dim"], float(self.cfg["scale_factor"])
        self._h = C.c_void_p()
        self._lib = L.lib()
        L.check(self._lib.ltt_vae_create(C.byref(c), self.device.index, C.byref(self._h)), "ltt_vae_create")
        self._finalized = False

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """autoencoder.load_state_dict(saved_ckpt['autoencoder']) (reference txt2img.py:107); encoder keys are skipped."""
        for k, v in sd.items():
            if k.startswith("decoder.") or k.startswith("post_quant_conv."):
                t = v.detach().to(torch.float32).contiguous()
                shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
                L.check(self._lib.ltt_vae_load_param(self._h, k.encode(), L.ptr(t), shape, t.dim(), 0 if t.is_cuda else 1),
                        f"ltt_vae_load_param({k})")
        self._finalized = False

    def finalize(self) -> None:
        L.check(self._lib.ltt_vae_finalize(self._h), "ltt_vae_finalize")
        self._finalized = True

    def _z(self, z):
        zz = z.detach().to(device=self.device, dtype=torch.float32).contiguous()
        if zz.dim() != 4 or zz.shape[1] != self.cfg["z_channels"]:
            raise L.LttError(f"VaeDecoder: latents must be [B,{self.cfg['z_channels']},h,w], got {tuple(zz.shape)}")
        if not self._finalized:
            self.finalize()
        return zz

    def decode(self, z: torch.Tensor, images_u8: bool = False):
        """z [B,4,h,w] -> img [B,3,8h,8w] fp32 (AutoencoderKL.decode); with images_u8 also the [B,8h,8w,3] uint8 images."""
        zz = self._z(z)
        B, _, h, w = zz.shape
        up = 2 ** (len(self.cfg["ch_mult"]) - 1)
        img = torch.empty(B, self.cfg["out_ch"], h * up, w * up, device=self.device)
        u8 = torch.empty(B, h * up, w * up, self.cfg["out_ch"], device=self.device, dtype=torch.uint8) if images_u8 else None
        with torch.cuda.device(self.device):
            L.check(self._lib.ltt_vae_decode(self._h, L.ptr(zz), B, h, w, L.ptr(img), L.ptr(u8), L.stream_ptr()), "ltt_vae_decode")
        self._keep = zz
        return (img, u8) if images_u8 else img

    def decode_to_uint8(self, z: torch.Tensor, out_host: Optional[torch.Tensor] = None, sync: bool = True) -> torch.Tensor:
        """z -> [B,8h,8w,3] uint8 HWC images on the HOST: decode with the clamp / scale / uint8 conversion fused into the
        last convolution, then ONE (pinned, asynchronous) device-to-host copy of the whole batch -- instead of the
        reference's per-sample `.cpu().numpy()` of fp32 CHW tensors (txt2img.py:320-323)."""
        zz = self._z(z)
        B, _, h, w = zz.shape
        up = 2 ** (len(self.cfg["ch_mult"]) - 1)
        u8 = torch.empty(B, h * up, w * up, self.cfg["out_ch"], device=self.device, dtype=torch.uint8)
        with torch.cuda.device(self.device):
            L.check(self._lib.ltt_vae_decode(self._h, L.ptr(zz), B, h, w, None, L.ptr(u8), L.stream_ptr()), "ltt_vae_decode")
        self._keep = zz
        if out_host is None:
            out_host = torch.empty(u8.shape, dtype=torch.uint8).pin_memory()
        out_host.copy_(u8, non_blocking=True)
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        return out_host

    @property
    def launch_count(self) -> int:
        return int(self._lib.ltt_vae_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ltt_vae_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
