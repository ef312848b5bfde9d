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


def gather_shards(local, world):
    """all_gather of equally-sized shards, concatenated in rank order (= global sample order)."""
    bufs = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(bufs, local.contiguous())
    return torch.cat(bufs, dim=0)
