"""Batch sharding of the sampler over the GPUs of one node.

Samples are independent through the whole PLMS loop (GroupNorm / LayerNorm / attention are per sample), so a global
batch is split into contiguous per-rank slices with no data-path collective; the only exchange is ONE all-gather of the
final latents (32 KB per 64x64 sample).  The reference has no multi-GPU inference (txt2img.py:535-536 hard-codes cuda:0).
Plain torch.distributed: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.distributed as dist

BATCH_KEYS = ("x", "context", "uc", "relations", "boxes", "masks", "text_embeddings")


def slice_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) rows of `rank`; the first `global_batch % world` ranks take one extra row."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Per-rank slice of every per-sample tensor of a sampler batch (keys of BATCH_KEYS that are present)."""
    B = next(v.shape[0] for k, v in batch.items() if k in BATCH_KEYS)
    lo, hi = slice_bounds(B, rank, world)
    out = {}
    for k, v in batch.items():
        if k in BATCH_KEYS:
            if v.shape[0] != B:
                raise ValueError(f"'{k}' has batch {v.shape[0]}, expected {B}")
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def gather_latents(z: torch.Tensor, global_batch: int) -> torch.Tensor:
    """All-gather the per-rank final latents into the [global_batch, ...] tensor, in rank order (one collective)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return z
    world = dist.get_world_size()
    sizes = [slice_bounds(global_batch, r, world) for r in range(world)]
    cap = max(hi - lo for lo, hi in sizes)
    pad = z
    if z.shape[0] < cap:       # ragged split: pad to the common size, trim after the gather
        pad = torch.cat([z, z.new_zeros((cap - z.shape[0],) + tuple(z.shape[1:]))])
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous())
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)])
