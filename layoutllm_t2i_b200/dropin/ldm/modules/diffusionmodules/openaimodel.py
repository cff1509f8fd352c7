"""UNetModel of the drop-in tree: same constructor, attributes, `state_dict` grammar and `forward(dict)` as the
reference (GLIGEN/ldm/modules/diffusionmodules/openaimodel.py:232-459), executed by the sm_100a library.

The `nn.Module` tree below is a *parameter container* (so `load_state_dict(saved_ckpt['model'])`, `.to(device)`,
`model.modules()` type scans and `.scale` writes keep working); the arithmetic of every layer is in
`libltt_b200.so`.  There is no PyTorch/CPU execution path: calling the model without a CUDA device raises.
"""
import os
import sys

import torch as th
import torch.nn as nn

from ldm.modules.attention import GatedSelfAttentionDense, SpatialTransformer
from ldm.modules.diffusionmodules.util import conv_nd, linear, normalization
from ldm.util import instantiate_from_config

_NO_STANDALONE = "{} executes inside UNetModel on the sm_100a engine (no stand-alone/CPU path)"


class TimestepEmbedSequential(nn.Sequential):
    """Block wrapper of the reference (:26-44); container only."""

    def forward(self, *a, **k):
        raise RuntimeError(_NO_STANDALONE.format("TimestepEmbedSequential"))


class Upsample(nn.Module):
    """nearest x2 + conv3x3 (reference :57-85) -> library: nearest-x2 copy kernel + the implicit-GEMM conv."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        if use_conv:
            self.conv = conv_nd(dims, channels, self.out_channels, 3, padding=padding)


class Downsample(nn.Module):
    """conv3x3 stride 2 (reference :88-114)."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        if not use_conv:
            raise NotImplementedError("conv_resample=False is not part of the LayoutLLM-T2I configuration")
        self.channels, self.out_channels, self.use_conv = channels, out_channels or channels, use_conv
        self.op = conv_nd(dims, channels, self.out_channels, 3, stride=2, padding=padding)


class ResBlock(nn.Module):
    """GN-SiLU-conv, + time embedding, GN-SiLU-conv, + skip (reference :117-231)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if use_scale_shift_norm or up or down or use_conv:
            raise NotImplementedError("ResBlock variant not used by the LayoutLLM-T2I configuration")
        self.channels, self.emb_channels, self.out_channels = channels, emb_channels, out_channels or channels
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(),
                                       conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)


def _find_sd_first_conv():
    """SD_input_conv_weight_bias.pth ships with the reference (GLIGEN/); openaimodel.py:393-398 reads it from there."""
    cands = [os.environ.get("LTT_SD_FIRST_CONV", "")]
    for base in [os.getcwd(), os.path.join(os.getcwd(), "GLIGEN")] + list(sys.path):
        cands.append(os.path.join(base or ".", "SD_input_conv_weight_bias.pth"))
    for c in cands:
        if c and os.path.isfile(c):
            return c
    raise FileNotFoundError("SD_input_conv_weight_bias.pth not found (set LTT_SD_FIRST_CONV or run from the "
                            "reference checkout); needed by restore_first_conv_from_SD")


class UNetModel(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, use_checkpoint=False, num_heads=8,
                 use_scale_shift_norm=False, transformer_depth=1, context_dim=None, fuser_type=None,
                 inpaint_mode=False, grounding_downsampler=None, grounding_tokenizer=None):
        super().__init__()
        if inpaint_mode or grounding_downsampler is not None:
            raise NotImplementedError("inpainting / grounding-downsampler UNets are outside the LayoutLLM-T2I hot path")
        assert fuser_type in ["gatedSA", "gatedSA2", "gatedCA"]
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions, self.dropout, self.channel_mult = attention_resolutions, dropout, channel_mult
        self.conv_resample, self.use_checkpoint, self.num_heads = conv_resample, use_checkpoint, num_heads
        self.context_dim, self.fuser_type, self.inpaint_mode = context_dim, fuser_type, inpaint_mode
        self.grounding_tokenizer_input = None          # set externally (reference txt2img.py:251,558)
        self.downsample_net = None
        self.additional_channel_from_downsampler = 0
        self.first_conv_type = "SD"
        self.first_conv_restorable = True

        emb_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, emb_dim), nn.SiLU(), linear(emb_dim, emb_dim))

        def res(cin, cout):
            return ResBlock(cin, emb_dim, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint)

        def st(ch):
            return SpatialTransformer(ch, key_dim=context_dim, value_dim=context_dim, n_heads=num_heads,
                                      d_head=ch // num_heads, depth=transformer_depth, fuser_type=fuser_type,
                                      use_checkpoint=use_checkpoint)

        # encoder: conv_in, then per level num_res_blocks x (ResBlock [+ SpatialTransformer]) and a Downsample
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        skip_chans, ch, ds = [model_channels], model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(st(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                skip_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(res(ch, ch), st(ch), res(ch, ch))
        # decoder: every block consumes cat([h, skip]) (reference :456)
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [res(ch + skip_chans.pop(), model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(st(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), nn.SiLU(), conv_nd(dims, model_channels, out_channels, 3, padding=1))
        self.position_net = instantiate_from_config(grounding_tokenizer)

        # engine state (not part of state_dict)
        self._engine = None
        self._engine_stale = True
        self._cond_key = None
        self._sd_conv = None
        self._register_load_state_dict_pre_hook(self._mark_stale)

    # ------------------------------------------------------------------------------------------------ engine plumbing
    def _mark_stale(self, *a, **k):
        # load_state_dict: the checkpoint's input_blocks.0.0 overwrites a swapped first conv, as in the reference
        self._engine_stale = True
        self._sd_conv_active = False

    def _param_version(self):
        # in-place edits (optimizer step, EMA swap, p.data.copy_) bump the tensors' version counters
        return sum(p._version for p in self.parameters())

    def _apply(self, fn, *a, **k):      # .to() / .cuda() / .half(): device copies change -> re-upload lazily
        self._engine_stale = True
        return super()._apply(fn, *a, **k)

    def engine_config(self):
        pn = self.position_net
        return dict(in_channels=self.in_channels, out_channels=self.out_channels, model_channels=self.model_channels,
                    attention_resolutions=list(self.attention_resolutions), num_res_blocks=self.num_res_blocks,
                    channel_mult=list(self.channel_mult), num_heads=self.num_heads, context_dim=self.context_dim,
                    grounding_in_dim=pn.in_dim, grounding_out_dim=pn.out_dim, fourier_freqs=pn.fourier_freqs,
                    max_objs=30)

    def engine(self, max_objs=30):
        """The library handle holding this model's weights (created / refreshed lazily)."""
        from layoutllm_t2i_b200.engine import Engine
        dev = self.out[2].weight.device
        if dev.type != "cuda":
            raise RuntimeError("UNetModel runs on a CUDA device only (sm_100a library, no CPU fallback): "
                               "call model.to('cuda') first")
        if self._engine is not None and (self._engine.device != dev or self._engine.cfg["max_objs"] != max_objs):
            self._engine.close()
            self._engine = None
        if self._engine is None:
            self._engine = Engine(dict(self.engine_config(), max_objs=max_objs), dev)
            self._engine_stale = True
        version = self._param_version()
        if self._engine_stale or version != getattr(self, "_engine_version", None):
            self._engine.load_state_dict(self.state_dict())
            self._engine.finalize()
            if self._sd_conv is not None and getattr(self, "_sd_conv_active", False):
                self._engine.set_first_conv(*self._sd_conv)
            else:
                self._engine.clear_first_conv()
            self._engine_stale = False
            self._engine_version = version
            self._cond_key = None
        return self._engine

    @staticmethod
    def _tkey(t):
        return None if t is None else (t.data_ptr(), t._version, tuple(t.shape), t.dtype)

    def _conditioning(self, eng, context, relations, grounding, n_grounded, H, W):
        key = (self._tkey(context), self._tkey(relations), n_grounded, H, W,
               None if grounding is None else tuple(self._tkey(grounding[k]) for k in ("boxes", "masks", "positive_embeddings")))
        if key != self._cond_key:
            eng.set_conditioning(context, relations, grounding, H, W, n_grounded)
            self._cond_key = key
            self._cond_refs = (context, relations, grounding)    # keep data_ptr keys valid

    def fuser_scale(self):
        """The gate multiplier set_alpha_scale wrote (reference txt2img.py:46-50): read from the first fuser."""
        for m in self.modules():
            if type(m) == GatedSelfAttentionDense:
                return float(m.scale)
        return 1.0

    # ------------------------------------------------------------------------------------------------ reference API
    def sd_first_conv(self):
        if self._sd_conv is None:
            sd = th.load(_find_sd_first_conv(), map_location="cpu")
            self._sd_conv = (sd["weight"].float(), sd["bias"].float())
        return self._sd_conv

    def set_sd_first_conv(self, weight, bias):
        """Extension: provide the Stable-Diffusion first-conv tensors directly instead of the .pth lookup."""
        self._sd_conv = (weight.detach().float().cpu(), bias.detach().float().cpu())

    def restore_first_conv_from_SD(self):
        """Swap input_blocks[0][0] for the Stable-Diffusion first conv (reference :393-408).  Permanent, as in the
        reference; the tensors are read from disk once and cached (the reference re-reads them 35x per image)."""
        if not self.first_conv_restorable:
            print("First conv layer is not restorable and skipped this process, probably because this is an inpainting model?")
            return
        if getattr(self, "_sd_conv_active", False):
            return
        w, b = self.sd_first_conv()
        conv = self.input_blocks[0][0]
        self.GLIGEN_first_conv_state_dict = {k: v.clone() for k, v in conv.state_dict().items()}
        with th.no_grad():
            conv.weight.copy_(w.to(conv.weight))
            conv.bias.copy_(b.to(conv.bias))
        self._sd_conv_active = True
        self.first_conv_type = "SD"
        if self._engine is not None and not self._engine_stale:
            self._engine.set_first_conv(w, b)
            self._engine_version = self._param_version()     # the in-place copy above is already in the engine

    def restore_first_conv_from_GLIGEN(self):
        raise NotImplementedError("not implemented in the reference either (openaimodel.py:410-411)")

    @th.no_grad()
    def forward(self, input):
        """eps = UNet(x, t | context, relations, grounding) (reference :413-459).  `input` keys: x [B,4,H,W],
        timesteps [B], context [B,77,768], relations [B,R,768], optional grounding_input {boxes, masks,
        positive_embeddings}; absent -> the null grounding input (reference :415-419)."""
        x = input["x"]
        B, _, H, W = x.shape
        grounding = input.get("grounding_input")
        max_objs = grounding["boxes"].shape[1] if grounding is not None else getattr(self.grounding_tokenizer_input, "max_box", 30)
        eng = self.engine(max_objs)
        self._conditioning(eng, input["context"], input["relations"], grounding, B if grounding is not None else 0, H, W)
        return eng.forward(x, input["timesteps"], self.fuser_scale())
