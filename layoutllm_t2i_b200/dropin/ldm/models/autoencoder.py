"""AutoencoderKL of the drop-in tree: the reference's constructor, `state_dict` grammar and `decode(z)` contract
(GLIGEN/ldm/models/autoencoder.py:17-44; Decoder / ResnetBlock / AttnBlock of ldm/modules/diffusionmodules/model.py),
with decode executed by the sm_100a library (`ltt_vae_*`).

The `nn.Module` tree below is a parameter container: `autoencoder.load_state_dict(saved_ckpt["autoencoder"])` is STRICT in
the callers (txt2img.py:107), so encoder and quant_conv keys exist too; the encoder itself is not on the inference path
(txt2img.py never calls `encode`) and is not implemented here.  Extension for callers that want it:
`decode_to_uint8(z)` returns the HWC uint8 images on the host with ONE pinned copy (the callers' clamp / *255 / uint8
per-sample loop, txt2img.py:320-323, fused into the decoder's last convolution)."""
import torch
import torch.nn as nn

_NO_STANDALONE = "{} executes inside AutoencoderKL.decode on the sm_100a engine (no stand-alone/CPU path)"


def _norm(c):
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)


class _Container(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(_NO_STANDALONE.format(type(self).__name__))


class ResnetBlock(_Container):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1, self.conv1 = _norm(in_channels), nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2, self.conv2 = _norm(out_channels), nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class AttnBlock(_Container):
    def __init__(self, c):
        super().__init__()
        self.norm = _norm(c)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1, 1, 0) for _ in range(4))


class _Resample(_Container):
    def __init__(self, c, stride, padding):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride, padding)


def _mid(c):
    m = nn.Module()
    m.block_1, m.attn_1, m.block_2 = ResnetBlock(c, c), AttnBlock(c), ResnetBlock(c, c)
    return m


class Encoder(_Container):
    """Parameter container with Encoder's key grammar (model.py:368-425)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, in_channels, resolution,
                 z_channels, double_z=True, **ignore):
        super().__init__()
        if attn_resolutions:
            raise NotImplementedError("attn_resolutions must be empty (LayoutLLM-T2I autoencoder)")
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        in_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i in range(len(ch_mult)):
            block_in, block_out = ch * in_mult[i], ch * ch_mult[i]
            d = nn.Module()
            d.block, d.attn = nn.ModuleList(), nn.ModuleList()
            for _ in range(num_res_blocks):
                d.block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            if i != len(ch_mult) - 1:
                d.downsample = _Resample(block_in, 2, 0)
            self.down.append(d)
        self.mid = _mid(block_in)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, 1, 1)


class Decoder(_Container):
    """Parameter container with Decoder's key grammar (model.py:462-536)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, in_channels, resolution,
                 z_channels, give_pre_end=False, tanh_out=False, **ignore):
        super().__init__()
        if attn_resolutions or give_pre_end or tanh_out:
            raise NotImplementedError("decoder variant not used by the LayoutLLM-T2I autoencoder")
        block_in = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = _mid(block_in)
        self.up = nn.ModuleList()
        for i in reversed(range(len(ch_mult))):
            block_out = ch * ch_mult[i]
            u = nn.Module()
            u.block, u.attn = nn.ModuleList(), nn.ModuleList()
            for _ in range(num_res_blocks + 1):
                u.block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            if i != 0:
                u.upsample = _Resample(block_in, 1, 1)
            self.up.insert(0, u)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)


class AutoencoderKL(nn.Module):
    def __init__(self, ddconfig, embed_dim, scale_factor=1):
        super().__init__()
        assert ddconfig["double_z"]
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim, self.scale_factor = embed_dim, scale_factor
        self._ddconfig = dict(ddconfig)
        self._engine = None
        self._engine_stale = True
        self._register_load_state_dict_pre_hook(self._mark_stale)

    def _mark_stale(self, *a, **k):
        self._engine_stale = True

    def _apply(self, fn, *a, **k):
        self._engine_stale = True
        return super()._apply(fn, *a, **k)

    def engine(self):
        """The library handle holding the decoder weights (created / refreshed lazily)."""
        from layoutllm_t2i_b200.vae import VaeDecoder
        dev = self.post_quant_conv.weight.device
        if dev.type != "cuda":
            raise RuntimeError("AutoencoderKL.decode runs on a CUDA device only (sm_100a library, no CPU fallback)")
        version = sum(p._version for p in self.parameters())
        if self._engine is not None and self._engine.device != dev:
            self._engine.close()
            self._engine = None
        if self._engine is None:
            d = self._ddconfig
            self._engine = VaeDecoder(dict(ch=d["ch"], out_ch=d["out_ch"], ch_mult=list(d["ch_mult"]), num_res_blocks=d["num_res_blocks"],
                                           z_channels=d["z_channels"], embed_dim=self.embed_dim, scale_factor=float(self.scale_factor)), dev)
            self._engine_stale = True
        if self._engine_stale or version != getattr(self, "_engine_version", None):
            self._engine.load_state_dict(self.state_dict())
            self._engine.finalize()
            self._engine_stale, self._engine_version = False, version
        return self._engine

    def encode(self, x):
        raise NotImplementedError("AutoencoderKL.encode is outside the LayoutLLM-T2I inference path (txt2img.py never "
                                  "encodes); use the reference's ldm.models.autoencoder for training / inpainting")

    @torch.no_grad()
    def decode(self, z):
        """dec = Decoder(post_quant_conv(z / scale_factor)) (reference :40-44) -> [B,3,8h,8w]."""
        return self.engine().decode(z).to(z.dtype if z.dtype in (torch.float16, torch.float32) else torch.float32)

    @torch.no_grad()
    def decode_to_uint8(self, z, out_host=None, sync=True):
        """[B,8h,8w,3] uint8 images on the host (the callers' post-processing, txt2img.py:320-323, fused + one pinned copy)."""
        return self.engine().decode_to_uint8(z, out_host, sync)
