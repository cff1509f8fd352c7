"""Deterministic stand-ins for the out-of-scope side models the reference callers instantiate from the checkpoint
config (CLIP text tower, VAE) and for the HF CLIP model/processor pair passed to generate_one_image."""
import hashlib
from types import SimpleNamespace

import torch


def _vec(text, n, dim=768):
    seed = int.from_bytes(hashlib.sha1(str(text).encode()).digest()[:8], "little") & 0x7FFFFFFFFFFFFFFF
    return torch.randn(n, dim, generator=torch.Generator().manual_seed(seed))


class TextEncoder(torch.nn.Module):
    """FrozenCLIPEmbedder's interface (ldm/modules/encoders/modules.py:144-184): encode(texts[, return_pooler_output])."""

    def __init__(self):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))
        self.device = "cpu"           # the callers probe `'device' in vars(m)` (txt2img.py:112-114)

    def encode(self, texts, return_pooler_output=False):
        dev = self.dummy.device
        seq = torch.stack([_vec(t, 77) for t in texts]).to(dev)
        if return_pooler_output:
            return seq, torch.cat([_vec("pool:" + t, 1) for t in texts]).to(dev)
        return seq


class Autoencoder(torch.nn.Module):
    """AutoencoderKL.decode stand-in: remembers the latents it was asked to decode and returns a 3-channel image."""

    def __init__(self):
        super().__init__()
        self.dummy = torch.nn.Parameter(torch.zeros(1))
        self.decoded = []

    def decode(self, z):
        self.decoded.append(z.detach().clone())
        return torch.tanh(torch.nn.functional.interpolate(z[:, :3].float(), scale_factor=8, mode="nearest"))


class ClipProcessor:
    def __call__(self, text=None, images=None, return_tensors="pt", padding=True):
        ids = torch.tensor([[int.from_bytes(hashlib.sha1(str(text).encode()).digest()[:4], "little") % 49408]])
        return dict(input_ids=ids, attention_mask=torch.ones_like(ids), _text=text)


class ClipModel:
    def __call__(self, input_ids=None, pixel_values=None, attention_mask=None, **kw):
        pooled = _vec("phrase:%d" % int(input_ids[0, 0]), 1).to(input_ids.device)
        return SimpleNamespace(text_model_output=SimpleNamespace(pooler_output=pooled))
