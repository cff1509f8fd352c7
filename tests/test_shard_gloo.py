"""N > 1 host logic on CPU: world_size-2 gloo processes shard a sampler batch and all-gather the 'latents'."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from layoutllm_t2i_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    batch = dict(x=torch.randn(B, 4, 8, 8, generator=g), context=torch.randn(B, 77, 16, generator=g),
                 boxes=torch.rand(B, 30, 4, generator=g), masks=torch.ones(B, 30), guidance=7.5)
    mine = shard.shard_batch(batch, rank, world)
    lo, hi = shard.slice_bounds(B, rank, world)
    assert mine["x"].shape[0] == hi - lo and mine["guidance"] == 7.5
    assert torch.equal(mine["boxes"], batch["boxes"][lo:hi])
    z = mine["x"] * 2 + rank * 0          # stand-in for the sampler: any per-sample function
    full = shard.gather_latents(z, B)
    assert torch.equal(full, batch["x"] * 2)
    torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 5])
def test_shard_and_gather_world2(tmp_path, B):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    assert torch.equal(a, b) and a.shape[0] == B


def test_slice_bounds_cover_batch():
    for B in (1, 7, 64):
        for world in (1, 2, 4, 8):
            spans = [shard.slice_bounds(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    with pytest.raises(ValueError):
        shard.slice_bounds(8, 2, 2)


def test_single_process_gather_is_identity():
    z = torch.randn(3, 4, 8, 8)
    assert shard.gather_latents(z, 3) is z
