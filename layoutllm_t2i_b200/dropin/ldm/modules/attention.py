"""Parameter containers with the reference's attention-module names and state_dict grammar
(reference GLIGEN/ldm/modules/attention.py).  They own weights and the gate attributes callers poke
(`.scale` on GatedSelfAttentionDense, reference txt2img.py:46-50); the arithmetic lives in the sm_100a library
(fused QKV GEMM -> tcgen05 flash attention -> out-proj, GEGLU GEMMs, box pooling), driven by UNetModel.
Names this file does not define (e.g. LinearAttention, used by the reference VAE) fall through to the reference."""
import torch
import torch.nn as nn

import _ltt_fallthrough

_NO_STANDALONE = "{} executes inside UNetModel on the sm_100a engine (no stand-alone/CPU path)"


class _Container(nn.Module):
    def forward(self, *args, **kwargs):
        raise RuntimeError(_NO_STANDALONE.format(type(self).__name__))


class GEGLU(_Container):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)      # rows [value | gate] (reference :44)


class FeedForward(_Container):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        inner = int(dim * mult)
        dim_out = dim if dim_out is None else dim_out
        first = GEGLU(dim, inner) if glu else nn.Sequential(nn.Linear(dim, inner), nn.GELU())
        self.net = nn.Sequential(first, nn.Dropout(dropout), nn.Linear(inner, dim_out))


class CrossAttention(_Container):
    def __init__(self, query_dim, key_dim, value_dim, heads=8, dim_head=64, dropout=0):
        super().__init__()
        inner = dim_head * heads
        self.scale, self.heads = dim_head ** -0.5, heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(key_dim, inner, bias=False)
        self.to_v = nn.Linear(value_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))


class SelfAttention(CrossAttention):
    def __init__(self, query_dim, heads=8, dim_head=64, dropout=0.):
        super().__init__(query_dim, query_dim, query_dim, heads, dim_head, dropout)


class _Gated(_Container):
    def _gates(self):
        self.register_parameter("alpha_attn", nn.Parameter(torch.tensor(0.)))
        self.register_parameter("alpha_dense", nn.Parameter(torch.tensor(0.)))
        self.scale = 1      # external multiplier of tanh(alpha) (set_alpha_scale)


class GatedSelfAttentionDense(_Gated):
    """GLIGEN fuser: x += scale*tanh(a)*SA(LN([x; linear(objs)]))[:, :N]; x += scale*tanh(d)*FF(LN(x)) (reference :204-234)."""

    def __init__(self, query_dim, context_dim, n_heads, d_head):
        super().__init__()
        self.linear = nn.Linear(context_dim, query_dim)
        self.attn = SelfAttention(query_dim=query_dim, heads=n_heads, dim_head=d_head)
        self.ff = FeedForward(query_dim, glu=True)
        self.norm1, self.norm2 = nn.LayerNorm(query_dim), nn.LayerNorm(query_dim)
        self._gates()


class GatedCrossAttentionDense(_Gated):
    """Kept for the callers' type checks (reference txt2img.py:47-50); the LayoutLLM-T2I checkpoint uses gatedSA."""

    def __init__(self, query_dim, key_dim, value_dim, n_heads, d_head):
        super().__init__()
        self.attn = CrossAttention(query_dim, key_dim, value_dim, n_heads, d_head)
        self.ff = FeedForward(query_dim, glu=True)
        self.norm1, self.norm2 = nn.LayerNorm(query_dim), nn.LayerNorm(query_dim)
        self._gates()


class RelationCrossAttention(_Gated):
    """Relation-aware fusion: box mean-pool of LN3(x) -> 30-token cross-attention over the relation embeddings + GEGLU
    FF -> masked scatter back (reference :284-359); `scale` stays 1 (set_alpha_scale does not touch this type)."""

    def __init__(self, query_dim, key_dim, value_dim, n_heads, d_head):
        super().__init__()
        self.attn = CrossAttention(query_dim, key_dim, value_dim, n_heads, d_head)
        self.ff = FeedForward(query_dim, glu=True)
        self.norm1, self.norm2, self.norm3 = (nn.LayerNorm(query_dim) for _ in range(3))
        self._gates()


class BasicTransformerBlock(_Container):
    def __init__(self, query_dim, key_dim, value_dim, n_heads, d_head, fuser_type, use_checkpoint=True):
        super().__init__()
        if fuser_type != "gatedSA":
            raise NotImplementedError("the B200 path implements fuser_type='gatedSA' (the LayoutLLM-T2I checkpoint)")
        self.attn1 = SelfAttention(query_dim=query_dim, heads=n_heads, dim_head=d_head)
        self.ff = FeedForward(query_dim, glu=True)
        self.attn2 = CrossAttention(query_dim, key_dim, value_dim, n_heads, d_head)
        self.norm1, self.norm2, self.norm3 = (nn.LayerNorm(query_dim) for _ in range(3))
        self.use_checkpoint = use_checkpoint
        self.fuser = GatedSelfAttentionDense(query_dim, key_dim, n_heads, d_head)
        self.rela_fuse = RelationCrossAttention(query_dim, key_dim, value_dim, n_heads, d_head)


class SpatialTransformer(_Container):
    def __init__(self, in_channels, key_dim, value_dim, n_heads, d_head, depth=1, fuser_type=None, use_checkpoint=True):
        super().__init__()
        if depth != 1:
            raise NotImplementedError("transformer_depth must be 1")
        self.in_channels = in_channels
        query_dim = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, query_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(query_dim, key_dim, value_dim, n_heads, d_head, fuser_type, use_checkpoint)])
        self.proj_out = nn.Conv2d(query_dim, in_channels, kernel_size=1, stride=1, padding=0)


_ltt_fallthrough.install(__name__, globals())
