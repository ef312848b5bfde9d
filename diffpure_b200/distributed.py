"""Host-side plumbing for the data-parallel path: one process per GPU, weights broadcast once,
contiguous batch shards, no collective inside the purification loop, one all_gather of the purified images.

The reference instead re-replicates the whole module on every forward through nn.DataParallel
(eval_sde_adv.py:227-229). Works with NCCL (GPU tensors) and gloo (CPU tensors; used by the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [start, end) of `total` samples owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_state_dict(sd, src=0, device="cpu"):
    """One collective for all weights: flatten (fixed key order) -> broadcast -> unflatten. Non-src ranks only need
    tensors of the right shapes. Returns fp32 CPU tensors keyed like `sd`."""
    names = list(sd.keys())
    flat = torch.cat([sd[k].detach().float().reshape(-1) for k in names]).to(device)
    if dist.get_rank() != src:
        flat.zero_()
    dist.broadcast(flat, src=src)
    flat = flat.cpu()
    out, off = {}, 0
    for k in names:
        n = sd[k].numel()
        out[k] = flat[off:off + n].reshape(sd[k].shape).clone()
        off += n
    return out


def broadcast_blob(blob, src=0):
    """The data-parallel weight distribution: rank `src` packs and uploads the constants once, every other rank builds
    only the layout (`WeightBlob(program, device, upload=False)`) and receives the packed bytes with ONE broadcast."""
    blob.broadcast(src)
    return blob


def gather_shards(local, world, total=None):
    """all_gather of the per-rank shards, concatenated in rank order (= global sample order). Shards produced by
    `shard_range` differ by at most one sample: every rank pads to the largest shard and the padding is trimmed after the
    collective (`total` = global sample count; None = equal shards)."""
    local = local.contiguous()
    if total is None:
        bufs = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(bufs, local)
        return torch.cat(bufs, dim=0)
    sizes = [shard_range(total, r, world) for r in range(world)]
    sizes = [b - a for a, b in sizes]
    assert local.shape[0] == sizes[dist.get_rank()], (local.shape[0], sizes)
    big = max(sizes)
    if local.shape[0] < big:
        pad = torch.zeros((big - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    bufs = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(bufs, local)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)
