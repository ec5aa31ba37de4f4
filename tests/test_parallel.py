"""CPU tests of the N>1 host logic: world_size-2 gloo processes shard a batch, encode their slice and all-gather.
The encoder stand-in is the oracle (test infrastructure) -- the CUDA library cannot run here; what is under test is
shard boundaries (ragged batches included), gather order and equality with the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ml_fastvlm_b200 import parallel as par


def test_shard_bounds_cover_and_are_contiguous():
    for batch in (1, 2, 5, 8, 256, 257):
        for world in (1, 2, 3, 4, 8):
            spans = [par.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
                assert b0 == a1 and b0 >= a0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert par.shard_bounds(256, 3, 8) == (96, 128)          # rank r takes [r*256/G, (r+1)*256/G)
    with pytest.raises(ValueError):
        par.shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        images = torch.rand(batch, 3, 8, 8)                                  # same on every rank
        w = torch.randn(7, 3 * 8 * 8)

        def encode(x):                                                       # per-image, batch-independent stand-in
            return (x.reshape(x.shape[0], -1) @ w.t()).reshape(x.shape[0], 1, 7)

        full = encode(images)
        got = par.encode_images_sharded(encode, images, gather=True)
        local = par.encode_images_sharded(encode, images, gather=False)
        a, b = par.shard_bounds(batch, rank, world)
        ok = torch.allclose(got, full) and got.shape == full.shape and torch.allclose(local, full[a:b])
        ok = ok and torch.equal(par.shard_batch(images), images[a:b])
        torch.save({"ok": bool(ok)}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5])
def test_two_rank_shard_and_allgather_matches_single_process(tmp_path, batch):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))["ok"], f"rank {r}"
