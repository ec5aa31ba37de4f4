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
    """Run `encode_fn` (e.g. `lambda x: engine.forward(x)[1]`) on this rank's shard; optionally all-gather (NCCL pass)."""
    batch = images.shape[0]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    # decided identically on EVERY rank before any collective: a rank-local raise would leave the others inside the
    # all-gather forever
    if gather and batch < world:
        raise ValueError(f"all-gather needs at least one image per rank (batch {batch} < world {world})")
    a, b = shard_bounds(batch, rank, world)
    local = encode_fn(images[a:b]) if b > a else None
    if not gather:
        return local
    return all_gather_tokens(local, batch, group)


def gather_slots(batch, world):
    """[(first_image, n_images)] of every rank's slot in the gathered [batch, N, H] buffer."""
    return [(a, b - a) for a, b in (shard_bounds(batch, r, world) for r in range(world))]


class GatheredEncoder:
    """Batch-sharded encode_images whose projector epilogue IS the all-gather.

    Every rank owns a symmetric-memory buffer [batch, N, H] (torch.distributed._symmetric_memory: CUDA-IPC / fabric handles
    exchanged once at construction; NVLink peer mappings).  `encode(images_of_my_shard)` makes ONE library call
    (`fvhd_forward_gather`): the projector GEMM's epilogue stores each output vector into this rank's slot of its own
    buffer and, in the same kernel, into the same slot of every peer's buffer.  A symmetric-memory barrier then orders
    all ranks' stores before any reader.  No NCCL collective is on the data path.
    """

    def __init__(self, engine, batch, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.engine = engine
        self.batch = int(batch)
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.batch < self.world:
            raise ValueError(f"gathered encode needs at least one image per rank (batch {self.batch} < world {self.world})")
        if self.world - 1 > 8:
            raise ValueError("at most 8 peers (one NVSwitch domain)")
        n, h = engine.num_tokens, engine.hidden
        self.buf = symm_mem.empty((self.batch, n, h), dtype=torch.bfloat16, device=engine.device)
        try:        # older torch needs the group enabled explicitly; newer versions do it inside rendezvous
            symm_mem.enable_symm_mem_for_group(self.group.group_name)
        except Exception:  # noqa: BLE001
            pass
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.slots = gather_slots(self.batch, self.world)
        a, nb = self.slots[self.rank]
        off = a * n * h * 2                      # byte offset of this rank's slot -- the same in every rank's buffer
        self.local = self.buf[a:a + nb]
        self.peer_ptrs = [int(self.hdl.buffer_ptrs[r]) + off for r in range(self.world) if r != self.rank]

    def encode(self, my_images, barrier=True):
        """my_images: this rank's [b_r,3,R,R] shard.  Returns the gathered [batch, N, H] tensor (valid after the barrier)."""
        a, nb = self.slots[self.rank]
        if my_images.shape[0] != nb:
            raise ValueError(f"rank {self.rank}: expected {nb} local images, got {my_images.shape[0]}")
        self.engine.forward_gather(my_images, self.local, self.peer_ptrs)
        if barrier:
            self.hdl.barrier(channel=0)          # stream-ordered: all ranks' peer stores precede the readers
        return self.buf
