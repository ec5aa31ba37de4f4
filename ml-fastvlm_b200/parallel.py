"""Multi-GPU plumbing for the encode_images path: batch sharding + (optional) token all-gather.

Images are independent through tower and projector (eval-mode BatchNorm, SE pools per image), so the path
shards by batch with replicated weights and NO collective inside it.  One process per GPU;
`torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used only when a single LLM
prefill needs the whole visual-token batch: one all-gather of the projected tokens.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, rank, world):
    """Contiguous split; the first `batch % world` ranks take one extra image.  -> (start, stop)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(images, rank=None, world=None):
    """This rank's slice of a [B,3,R,R] batch (or of a list of images)."""
    if rank is None:
        rank, world = dist.get_rank(), dist.get_world_size()
    n = len(images) if isinstance(images, (list, tuple)) else images.shape[0]
    a, b = shard_bounds(n, rank, world)
    return images[a:b]


def all_gather_tokens(local_tokens, batch, group=None):
    """[b_r, N, H] per rank -> [batch, N, H] on every rank, in global image order.
    Ragged shards (batch % world != 0) are padded to the largest shard for the collective and trimmed after."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(batch, r, world) for r in range(world)]
    bmax = max(b - a for a, b in sizes)
    n, h = local_tokens.shape[1], local_tokens.shape[2]
    a, b = sizes[rank]
    if local_tokens.shape[0] != b - a:
        raise ValueError(f"rank {rank}: expected {b - a} local images, got {local_tokens.shape[0]}")
    send = local_tokens
    if b - a < bmax:
        pad = torch.zeros(bmax - (b - a), n, h, dtype=local_tokens.dtype, device=local_tokens.device)
        send = torch.cat([local_tokens, pad], 0)
    out = torch.empty(world * bmax, n, h, dtype=local_tokens.dtype, device=local_tokens.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    parts = [out[r * bmax: r * bmax + (sizes[r][1] - sizes[r][0])] for r in range(world)]
    return torch.cat(parts, 0)


def encode_images_sharded(encode_fn, images, gather=True, group=None):
    """Run `encode_fn` (e.g. `lambda x: engine.forward(x)[1]`) on this rank's shard; optionally all-gather."""
    batch = images.shape[0]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    a, b = shard_bounds(batch, rank, world)
    local = encode_fn(images[a:b]) if b > a else None
    if not gather:
        return local
    if local is None:
        raise ValueError("all-gather needs at least one image per rank")
    return all_gather_tokens(local, batch, group)
