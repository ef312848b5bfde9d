"""CPU: the N>1 host logic on a world_size-2 gloo group -- weight broadcast, shard ranges, gather order."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffpure_b200 import distributed as D


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(123)
        ref = {"a.weight": torch.randn(7, 5, generator=g), "b.bias": torch.randn(11, generator=g),
               "c.W": torch.randn(3, 3, 2, generator=g)}
        mine = ref if rank == 0 else {k: torch.full_like(v, float("nan")) for k, v in ref.items()}
        got = D.broadcast_state_dict(mine, src=0)
        ok = all(torch.equal(got[k], ref[k]) for k in ref) and list(got) == list(ref)
        s, e = D.shard_range(10, rank, world)
        full = torch.arange(10 * 3, dtype=torch.float32).reshape(10, 3)
        gathered = D.gather_shards(full[s:e], world)          # equal shards: 5 + 5
        ok = ok and torch.equal(gathered, full)
        s, e = D.shard_range(11, rank, world)                 # ragged shards: 6 + 5, padded for the collective
        full11 = torch.arange(11 * 3, dtype=torch.float32).reshape(11, 3)
        ok = ok and torch.equal(D.gather_shards(full11[s:e], world, total=11), full11)
        # packed weight blob: rank 0 packs the real constants, rank 1 only the layout; ONE broadcast of the bytes
        from diffpure_b200 import lowering_ncsnpp as L, synthetic
        from diffpure_b200.engine import WeightBlob
        from types import SimpleNamespace
        cfg = SimpleNamespace(image_size=16, num_channels=3, nf=64, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(8,))
        shapes = L.param_shapes(cfg)
        real = synthetic.random_state_dict(shapes, seed=4)
        sd = real if rank == 0 else {k: torch.zeros(v) for k, v in shapes.items()}
        prog = L.lower(cfg, sd, 2)
        blob = D.broadcast_blob(WeightBlob(prog, "cpu", upload=(rank == 0)), src=0)
        want = L.lower(cfg, real, 2)
        for t in want.tensors:
            if t.init is not None:
                ref_t = t.init.to(torch.bfloat16).float() if t.dtype == "bf16" else t.init.float()
                ok = ok and torch.equal(blob.tensor(t.name, t.dtype), ref_t.reshape(-1))
        ok = ok and blob.matches(L.lower(cfg, real, 5))       # the layout does not depend on the batch size
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_broadcast_shard_gather_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + os.getpid() % 200
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_ranges_partition():
    for total in (1, 7, 64, 512, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
