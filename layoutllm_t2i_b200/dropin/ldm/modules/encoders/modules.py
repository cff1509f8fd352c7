"""FrozenCLIPEmbedder of the drop-in tree: the reference's constructor, `state_dict` grammar and
`encode` / `encode_one_token` contracts (GLIGEN/ldm/modules/encoders/modules.py:144-182) with the text transformer
executed by the sm_100a library (`ltt_clip_*`).

The `nn.Module` tree is a parameter container in the grammar of transformers' CLIPTextModel
(`transformer.text_model.embeddings.token_embedding.weight`, ...), so `text_encoder.load_state_dict(saved_ckpt
["text_encoder"])` (txt2img.py:108) works unchanged; the persistent `position_ids` buffer older transformers
versions saved is accepted.  Tokenisation stays with transformers' CLIPTokenizer (host side, third-party): it is loaded
on first use, or handed in with `set_tokenizer` (any callable with CLIPTokenizer's call signature).

Extension for callers: `encode_many(texts)` returns (last_hidden_state, pooler_output) of ANY number of strings in one
pass -- `layoutllm_t2i_b200.clip.prepare_conditioning` builds the whole conditioning of an image on it.
The other encoders of the reference module (BERT, class / spatial embedders, CLIP image embedder) are not on the
LayoutLLM-T2I path and are not provided."""
import torch
import torch.nn as nn

_NO_STANDALONE = "{} executes inside FrozenCLIPEmbedder on the sm_100a engine (no stand-alone/CPU path)"


class AbstractEncoder(nn.Module):
    def encode(self, *args, **kwargs):
        raise NotImplementedError


class _Container(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(_NO_STANDALONE.format(type(self).__name__))


class _Attention(_Container):
    def __init__(self, w):
        super().__init__()
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (nn.Linear(w, w) for _ in range(4))


class _MLP(_Container):
    def __init__(self, w, f):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(w, f), nn.Linear(f, w)


class _Layer(_Container):
    def __init__(self, w, f, eps):
        super().__init__()
        self.self_attn = _Attention(w)
        self.layer_norm1 = nn.LayerNorm(w, eps=eps)
        self.mlp = _MLP(w, f)
        self.layer_norm2 = nn.LayerNorm(w, eps=eps)


class _TextModel(_Container):
    def __init__(self, cfg):
        super().__init__()
        w, f, eps = cfg["hidden_size"], cfg["intermediate_size"], cfg["layer_norm_eps"]
        self.embeddings = nn.Module()
        self.embeddings.token_embedding = nn.Embedding(cfg["vocab_size"], w)
        self.embeddings.position_embedding = nn.Embedding(cfg["max_position_embeddings"], w)
        self.encoder = nn.Module()
        self.encoder.layers = nn.ModuleList([_Layer(w, f, eps) for _ in range(cfg["num_hidden_layers"])])
        self.final_layer_norm = nn.LayerNorm(w, eps=eps)


class _Transformer(_Container):
    """Parameter container with CLIPTextModel's key grammar."""

    def __init__(self, cfg):
        super().__init__()
        self.text_model = _TextModel(cfg)


class FrozenCLIPEmbedder(AbstractEncoder):
    """Uses the CLIP transformer encoder for text (reference modules.py:144-182)."""

    def __init__(self, version="openai/clip-vit-large-patch14", device="cuda", max_length=77, text_config=None):
        super().__init__()
        from layoutllm_t2i_b200.clip import default_clip_text_config
        self.version = version
        self._cfg = dict(default_clip_text_config(), **(text_config or {}))
        self._tokenizer = None
        self.transformer = _Transformer(self._cfg)
        self.device = device
        self.max_length = max_length
        self._engine = None
        self._engine_stale = True
        self._register_load_state_dict_pre_hook(self._pre_load)
        self.freeze()

    # ---- reference surface
    @property
    def tokenizer(self):
        if self._tokenizer is None:
            from transformers import CLIPTokenizer
            self._tokenizer = CLIPTokenizer.from_pretrained(self.version)
        return self._tokenizer

    def set_tokenizer(self, tok):
        self._tokenizer = tok

    def freeze(self):
        self.transformer = self.transformer.eval()
        for param in self.parameters():
            param.requires_grad = False

    @torch.no_grad()
    def forward(self, text, return_pooler_output=False):
        batch_encoding = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True,
                                        return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        z, pooled = self.engine().encode_ids(batch_encoding["input_ids"])
        return (z, pooled) if return_pooler_output else z

    def encode(self, text, return_pooler_output=False):
        return self(text, return_pooler_output)

    @torch.no_grad()
    def encode_one_token(self, text, return_pooler_output=True):
        inputs = self.tokenizer(text=text, return_tensors="pt")
        z, pooled = self.engine().encode_ids(inputs["input_ids"])
        return pooled if return_pooler_output else z

    # ---- extension
    @torch.no_grad()
    def encode_many(self, texts):
        """(last_hidden_state [n, 77, W], pooler_output [n, W]) of n strings in one pass."""
        return self.forward(list(texts), return_pooler_output=True)

    def tokenize_padded(self, texts):
        """[n, max_length] ids exactly as `forward` feeds them (for layoutllm_t2i_b200.clip.prepare_conditioning)."""
        return self.tokenizer(list(texts), truncation=True, max_length=self.max_length, return_length=True,
                              return_overflowing_tokens=False, padding="max_length", return_tensors="pt")["input_ids"]

    # ---- engine plumbing
    def _pre_load(self, state_dict, prefix, *a, **k):
        self._engine_stale = True
        state_dict.pop(prefix + "transformer.text_model.embeddings.position_ids", None)   # buffer of older transformers

    def _apply(self, fn, *a, **k):
        self._engine_stale = True
        return super()._apply(fn, *a, **k)

    def engine(self):
        """The library handle holding the text tower (created / refreshed lazily)."""
        from layoutllm_t2i_b200.clip import ClipTextEncoder
        dev = self.transformer.text_model.final_layer_norm.weight.device
        if dev.type != "cuda":
            raise RuntimeError("FrozenCLIPEmbedder runs on a CUDA device only (sm_100a library, no CPU fallback)")
        version = sum(p._version for p in self.parameters())
        if self._engine is not None and self._engine.device != dev:
            self._engine.close()
            self._engine = None
        if self._engine is None:
            self._engine = ClipTextEncoder(self._cfg, dev)
            self._engine_stale = True
        if self._engine_stale or version != getattr(self, "_engine_version", None):
            self._engine.load_state_dict(self.state_dict())
            self._engine.finalize()
            self._engine_stale, self._engine_version = False, version
        return self._engine


def __getattr__(name):
    raise AttributeError(f"ldm.modules.encoders.modules.{name} is not part of the LayoutLLM-T2I inference path; the drop-in "
                         "tree provides FrozenCLIPEmbedder only (use the reference module for the other encoders)")
