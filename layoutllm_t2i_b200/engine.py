"""Python handle on the sm_100a UNet / PLMS engine (``ltt_model`` of include/ltt_b200.h).

Holds no arithmetic: every call forwards PyTorch-owned device pointers to the C-ABI.  The drop-in ``ldm`` module
tree (``layoutllm_t2i_b200/dropin``) and bench.py use this class; tests call it directly.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib as L


def default_unet_config() -> dict:
    """Hyper-parameters of the LayoutLLM-T2I UNet (reference GLIGEN/configs/coco2014.yaml:8-30)."""
    return dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, context_dim=768,
                grounding_in_dim=768, grounding_out_dim=768, fourier_freqs=8, max_objs=30)


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class Engine:
    """One engine per (UNet, CUDA device).  Not thread safe; stream ordered on the current torch stream."""

    def __init__(self, cfg: dict, device: int | torch.device | str = 0):
        if not torch.cuda.is_available():
            raise L.LttError("layoutllm_t2i_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.cfg = dict(default_unet_config(), **{k: v for k, v in cfg.items() if k in default_unet_config()})
        c = L.UNetConfig()
        c.in_channels, c.out_channels = self.cfg["in_channels"], self.cfg["out_channels"]
        c.model_channels, c.num_res_blocks = self.cfg["model_channels"], self.cfg["num_res_blocks"]
        mult, ar = list(self.cfg["channel_mult"]), list(self.cfg["attention_resolutions"])
        c.n_levels, c.n_attn_res = len(mult), len(ar)
        for i, v in enumerate(mult):
            c.channel_mult[i] = int(v)
        for i, v in enumerate(ar):
            c.attention_resolutions[i] = int(v)
        c.num_heads, c.context_dim = self.cfg["num_heads"], self.cfg["context_dim"]
        c.grounding_in_dim, c.grounding_out_dim = self.cfg["grounding_in_dim"], self.cfg["grounding_out_dim"]
        c.fourier_freqs, c.max_objs = self.cfg["fourier_freqs"], self.cfg["max_objs"]
        self._h = C.c_void_p()
        self._lib = L.lib()
        L.check(self._lib.ltt_create(C.byref(c), self.device.index, C.byref(self._h)), "ltt_create")
        self._finalized = False
        self._cond_key = None
        self._keep = []          # tensors whose storage the library may still be reading (stream ordered)

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """model.load_state_dict(saved_ckpt['model']) (reference txt2img.py:106): reference key names, fp32."""
        for k, v in sd.items():
            self.load_param(k, v)

    def load_param(self, key: str, v: torch.Tensor) -> None:
        t = v.detach().to(torch.float32).contiguous()
        shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
        L.check(self._lib.ltt_load_param(self._h, key.encode(), L.ptr(t), shape, t.dim(), 0 if t.is_cuda else 1),
                f"ltt_load_param({key})")
        self._finalized = False

    def finalize(self) -> None:
        L.check(self._lib.ltt_finalize(self._h), "ltt_finalize")
        self._finalized = True
        self._cond_key = None

    def set_first_conv(self, weight: torch.Tensor, bias: torch.Tensor) -> None:
        """UNetModel.restore_first_conv_from_SD (reference openaimodel.py:393-405); permanent."""
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        L.check(self._lib.ltt_set_first_conv(self._h, L.ptr(w), L.ptr(b), 0 if w.is_cuda else 1), "ltt_set_first_conv")

    def clear_first_conv(self) -> None:
        """Back to the checkpoint's own input_blocks.0.0 (what a later model.load_state_dict does in the reference)."""
        L.check(self._lib.ltt_clear_first_conv(self._h), "ltt_clear_first_conv")

    # ------------------------------------------------------------------ per-image conditioning
    def set_conditioning(self, context: torch.Tensor, relations: torch.Tensor, grounding: Optional[dict],
                         H: int, W: int, n_grounded: Optional[int] = None) -> None:
        """Everything UNetModel.forward derives from its non-x inputs (reference openaimodel.py:413-446).

        ``grounding`` = {boxes [Bg,30,4], masks [Bg,30], positive_embeddings [Bg,30,768]} for the first ``Bg`` batch
        rows (default: all rows when given, none when None); the rest use the null grounding input.
        """
        if not self._finalized:
            self.finalize()
        dev = self.device
        ctx, rel = _f32c(context, dev), _f32c(relations, dev)
        B = ctx.shape[0]
        if grounding is None:
            ng, boxes, masks, emb = 0, None, None, None
        else:
            boxes, masks = _f32c(grounding["boxes"], dev), _f32c(grounding["masks"], dev)
            emb = _f32c(grounding["positive_embeddings"], dev)
            ng = boxes.shape[0] if n_grounded is None else n_grounded
        with torch.cuda.device(dev):
            L.check(self._lib.ltt_set_conditioning(self._h, L.ptr(ctx), ctx.shape[1], L.ptr(rel), rel.shape[1],
                                                   L.ptr(boxes), L.ptr(masks), L.ptr(emb), B, ng, H, W,
                                                   L.stream_ptr()), "ltt_set_conditioning")
        self._keep = [ctx, rel, boxes, masks, emb]
        self.B, self.H, self.W = B, H, W

    # ------------------------------------------------------------------ compute
    def forward(self, x: torch.Tensor, timesteps: torch.Tensor, alpha_scale: float = 1.0) -> torch.Tensor:
        """UNetModel.forward (reference openaimodel.py:413-459) on the cached conditioning -> eps [B,4,H,W] fp32."""
        xx = _f32c(x, self.device)
        tt = _f32c(timesteps, self.device).reshape(-1)
        if xx.dim() != 4 or xx.shape[1] != self.cfg["in_channels"] or tt.numel() != xx.shape[0]:
            raise L.LttError(f"Engine.forward: x {tuple(xx.shape)} / timesteps {tuple(tt.shape)} do not form a "
                             f"[B,{self.cfg['in_channels']},H,W] batch with one timestep per sample")
        out = torch.empty(xx.shape[0], self.cfg["out_channels"], xx.shape[2], xx.shape[3], device=self.device)
        with torch.cuda.device(self.device):      # the library checks (B, H, W) against the cached conditioning
            L.check(self._lib.ltt_unet_forward(self._h, L.ptr(xx), L.ptr(tt), xx.shape[0], xx.shape[2], xx.shape[3],
                                               float(alpha_scale), L.ptr(out), L.stream_ptr()), "ltt_unet_forward")
        self._keep_fw = (xx, tt)
        return out

    def plms_sample(self, x: torch.Tensor, timesteps: Sequence[int], alphas, alphas_prev, sqrt_1m_alphas,
                    alpha_sched: Optional[Sequence[float]], guidance: float,
                    sd_first_conv: Optional[tuple] = None) -> torch.Tensor:
        """PLMSSampler.plms_sampling (reference plms.py:64-163): x [Bimg,4,H,W] start noise -> final latent."""
        S = len(timesteps)
        xx = _f32c(x, self.device).clone()
        ts = (C.c_int * S)(*[int(t) for t in timesteps])
        fa = lambda a: (C.c_float * S)(*[float(v) for v in np.asarray(a, dtype=np.float64)])
        sched = fa(alpha_sched) if alpha_sched is not None else None
        w = b = None
        if sd_first_conv is not None:
            w, b = _f32c(sd_first_conv[0], self.device), _f32c(sd_first_conv[1], self.device)
        with torch.cuda.device(self.device):
            L.check(self._lib.ltt_plms_sample(self._h, L.ptr(xx), xx.shape[0], S, ts, fa(alphas), fa(alphas_prev),
                                              fa(sqrt_1m_alphas), sched, float(guidance), L.ptr(w), L.ptr(b),
                                              L.stream_ptr()), "ltt_plms_sample")
        return xx

    PROFILE_CLASSES = ("gemm_tc", "attn_tc", "groupnorm", "layernorm", "forward", "relation")

    def profile(self, mode) -> None:
        """0 / False: off.  1 / True: eager launches, every class launch bracketed by CUDA events.  2: the brackets are
        event-record nodes inside the replayed CUDA graphs (calling it again with 2 restarts the replay counts)."""
        L.check(self._lib.ltt_profile_enable(self._h, int(mode)), "ltt_profile_enable")

    def profile_report(self) -> dict:
        """{class: {ms, flops, bytes, launches}} of the launches since profile(True) (CUDA events on the launch stream)."""
        out = {}
        for i, name in enumerate(self.PROFILE_CLASSES):
            ms, fl, by, n = C.c_double(), C.c_double(), C.c_double(), C.c_int64()
            L.check(self._lib.ltt_profile_report(self._h, i, C.byref(ms), C.byref(fl), C.byref(by), C.byref(n)), "ltt_profile_report")
            out[name] = dict(ms=ms.value, flops=fl.value, bytes=by.value, launches=n.value)
        return out

    def forward_with_taps(self, x, timesteps, alpha_scale=1.0, capacity_elems=1 << 28):
        """Debug: forward + {name: fp32 [rows, cols]} of the intermediate activations (parity tests)."""
        buf = torch.empty(capacity_elems, device=self.device)
        L.check(self._lib.ltt_debug_set_taps(self._h, L.ptr(buf), capacity_elems), "ltt_debug_set_taps")
        try:
            out = self.forward(x, timesteps, alpha_scale)
            torch.cuda.synchronize(self.device)
            taps = {}
            name = C.create_string_buffer(256)
            off, rows, cols = C.c_int64(), C.c_int64(), C.c_int64()
            for i in range(self._lib.ltt_debug_tap_count(self._h)):
                self._lib.ltt_debug_tap_info(self._h, i, name, 256, C.byref(off), C.byref(rows), C.byref(cols))
                taps[name.value.decode()] = buf[off.value: off.value + rows.value * cols.value].view(rows.value, cols.value).clone()
        finally:
            self._lib.ltt_debug_set_taps(self._h, None, 0)
        return out, taps

    @property
    def launch_count(self) -> int:
        return int(self._lib.ltt_launch_count(self._h))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ltt_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
