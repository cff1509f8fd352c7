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


class HashTokenizer:
    """CLIPTokenizer's call surface on a deterministic word hash: <bos> word ids <eos>, padded with <eos> (as the
    clip-vit-large-patch14 tokenizer does) to `max_length` when padding == "max_length", else to the longest row."""

    def __init__(self, vocab_size=49408):
        self.vocab, self.bos, self.eos = vocab_size, vocab_size - 2, vocab_size - 1

    def _row(self, text):
        words = [int.from_bytes(hashlib.sha1(w.encode()).digest()[:4], "little") % (self.vocab - 3) + 1 for w in str(text).split()]
        return [self.bos] + words + [self.eos]

    def __call__(self, text=None, truncation=False, max_length=77, padding=False, return_tensors="pt", **kw):
        texts = [text] if isinstance(text, str) else list(text)
        rows = [self._row(t) for t in texts]
        if truncation:
            rows = [r[:max_length - 1] + [self.eos] if len(r) > max_length else r for r in rows]
        width = max_length if padding == "max_length" else max(len(r) for r in rows)
        ids = torch.tensor([r + [self.eos] * (width - len(r)) for r in rows])
        mask = torch.tensor([[1] * len(r) + [0] * (width - len(r)) for r in rows])
        return dict(input_ids=ids, attention_mask=mask)


class HashProcessor:
    """CLIPProcessor's text call surface (txt2img.py:148: processor(text=..., return_tensors="pt", padding=True)) on HashTokenizer."""

    def __init__(self, vocab_size=49408):
        self.tok = HashTokenizer(vocab_size)

    def __call__(self, text=None, images=None, return_tensors="pt", padding=True):
        return self.tok(text, padding=padding)
