"""Seeded random-init weights for benchmarking without checkpoints (no network, no pretrained files).

Every tensor -- including the layers the reference zero-initialises -- gets variance-preserving values so the
benchmarked arithmetic is representative (no exact zeros short-circuiting anything)."""
import zlib

import torch


def random_state_dict(shapes, seed=0):
    sd = {}
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
        shape = tuple(shape)
        leaf = name.rsplit(".", 1)[-1]
        if len(shape) == 1:
            sd[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if leaf == "weight" else 0.1 * torch.randn(shape, generator=g)
            continue
        fan_in = shape[0] if leaf == "W" else int(torch.tensor(shape[1:]).prod())
        sd[name] = (torch.rand(shape, generator=g) * 2.0 - 1.0) * (3.0 / fan_in) ** 0.5
    return sd
