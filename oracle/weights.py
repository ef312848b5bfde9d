"""Deterministic weight factory -- test infrastructure only.

No pretrained checkpoints exist offline, and the reference zero-initialises every res-block's second
conv, every attention output projection and the output conv (score_sde/models/layers.py:88-91 via
init_scale=0; guided_diffusion/unet.py:218-220,302,623 via zero_module), which would make random-init
parity vacuous. This factory therefore fills *every* tensor with seeded,
variance-preserving values keyed by the parameter's name, so the same weights are reproduced on any
machine with the same torch build (the GPU box runs the same image).
"""
import zlib

import torch


def make_tensor(name, shape, seed):
    g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    is_norm = ("GroupNorm" in name) or ("norm" in name.lower() and len(shape) == 1) or \
              (len(shape) == 1 and leaf == "weight")
    if len(shape) == 1:
        if leaf in ("weight",) or (is_norm and leaf == "weight"):
            return 1.0 + 0.1 * torch.randn(shape, generator=g)      # norm scale
        return 0.1 * torch.randn(shape, generator=g)                 # biases / norm shift
    # dense / conv / NIN weights: uniform with variance 1/fan_in (keeps activations O(1) through the net)
    if leaf == "W":                                                  # NIN: [in, out]
        fan_in = shape[0]
    else:                                                            # [out, in, (kh, kw)]
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
    bound = (3.0 / fan_in) ** 0.5
    return (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound


def make_state_dict(shapes, seed=0):
    """shapes: ordered mapping name -> shape. Returns name -> fp32 tensor."""
    return {k: make_tensor(k, v, seed) for k, v in shapes.items()}
