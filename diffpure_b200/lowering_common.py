"""Shared pieces of the UNet lowerings: weight packing and the stream-tensor record."""
import math
from dataclasses import dataclass
from typing import Optional

import torch

from .program import ASeg, Program, Tensor, stats_rows

INV_SQRT2 = 1.0 / math.sqrt(2.0)


@dataclass
class Act:
    """An fp32 NHWC residual-stream tensor with its per-channel GroupNorm partial statistics."""
    t: Tensor
    C: int
    H: int
    W: int
    stats: Optional[Tensor]
    P: int  # partials per sample


def pack_conv3x3(w):
    """[Cout, Cin, 3, 3] -> [Cout, 9*Cin] with K index = (ky*3+kx)*Cin + ci (tap-major, matches the TMA tap loop)."""
    co, ci = w.shape[0], w.shape[1]
    return w.permute(0, 2, 3, 1).reshape(co, 9 * ci).contiguous()


def pack_conv1x1(w):
    """[Cout, Cin, 1, 1] or [Cout, Cin] -> [Cout, Cin]."""
    return w.reshape(w.shape[0], -1).contiguous()


def pack_conv_in(w):
    """[Cout, 3, 3, 3] -> [27, Cout] with row = (ky*3+kx)*3 + ci."""
    return w.permute(2, 3, 1, 0).reshape(27, w.shape[0]).contiguous()


def pack_conv_out(w):
    """[Cout, C, 3, 3] -> [9, C, Cout]."""
    return w.permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0]).contiguous()


def pad_rows(w, mult=128):
    """Zero-pad the leading (output) dimension to a multiple of `mult`."""
    n = w.shape[0]
    n2 = (n + mult - 1) // mult * mult
    if n2 == n:
        return w
    pad = torch.zeros((n2 - n,) + tuple(w.shape[1:]), dtype=w.dtype)
    return torch.cat([w, pad], dim=0)


def new_act(prog: Program, name, B, C, H, W, with_stats=True):
    t = prog.tensor(name, B * H * W * C, "f32")
    if with_stats:
        rows, P = stats_rows(B, H * W)
        st = prog.tensor(name + ".stats", rows * C * 2, "f32")
    else:
        st, P = None, 0
    return Act(t, C, H, W, st, P)


def act_seg(t, C, taps=1, stride=1, c_total=None, offset=0):
    from .program import view
    return ASeg(view(t, offset), C, C if c_total is None else c_total, taps, stride, 1 if (taps == 9 and stride == 1) else 0)
