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
    pre: Optional[Tensor] = None    # bf16 act(GroupNorm(x)) already produced for the next block by the producer's epilogue
    raw16: Optional[Tensor] = None  # bf16 copy of x written next to it (operand of the next block's 1x1 shortcut)


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


def lower_attention(prog, name, hn, wq, wk, wv, bq, bk, bv, B, T, C, heads, scale, rec=None):
    """Self-attention core on a normalised bf16 input hn [B*T, C]: returns o bf16 [B*T, C] with channel h*d + c.

    wq/wk/wv: [C, C] (out, in) with output rows ordered head-major; b*: [C].
      T <= 64        : one GEMM for q|k|v, then the whole-sequence smem kernel
      T in {128,256} : q|k GEMM, V^T GEMM (weights as A operand), tcgen05 S GEMM with the softmax epilogue
                       (numerator + row sums), tcgen05 O GEMM scaled by 1/rowsum
      longer T       : fp32 logits to HBM, row-softmax kernel, O GEMM
    """
    from .program import view
    d = C // heads
    o = prog.tensor(name + ".o", B * T * C, "bf16")
    if T <= 64:
        qkv = prog.tensor(name + ".qkv", B * T * 3 * C, "bf16")
        prog.gemm([act_seg(hn, C)], prog.const_bf16(name + ".wqkv", torch.cat([wq, wk, wv], 0)), 3 * C, C, 1, 1, B * T,
                  3 * C, bias=prog.const_f32(name + ".bqkv", torch.cat([bq, bk, bv])), out_bf16=qkv)
        prog.attn_small(qkv, o, B, T, heads, d, scale)
        if rec is not None:
            rec.update(qkv=qkv)
        return o
    assert T % 128 == 0 and d % 64 == 0, "tensor-core attention needs T % 128 == 0 and head dim % 64 == 0"
    qk = prog.tensor(name + ".qk", B * T * 2 * C, "bf16")
    prog.gemm([act_seg(hn, C)], prog.const_bf16(name + ".wqk", torch.cat([wq, wk], 0)), 2 * C, C, 1, 1, B * T, 2 * C,
              bias=prog.const_f32(name + ".bqk", torch.cat([bq, bk])), out_bf16=qk)
    # V^T per sample (all heads stacked): [C, T] = Wv[C, C] . hn_b[T, C]^T  (weights as the A operand, bias along M)
    vt = prog.tensor(name + ".vt", B * C * T, "bf16")
    prog.gemm([act_seg(prog.const_bf16(name + ".wv", wv), C)], hn, B * T, C, 1, 1, C, T, batch=B, a_batch_rows=0,
              b_batch_rows=T, out_batch_stride=C * T, bias=prog.const_f32(name + ".bv", bv), bias_along_m=1,
              out_bf16=vt, ldc=T)
    pm = prog.tensor(name + ".p", B * heads * T * T, "bf16")
    s_args = dict(batch=B * heads, inner=heads, a_batch_rows=T, a_inner_k=d, b_batch_rows=T, b_inner_k=d, w_cols=C,
                  out_batch_stride=heads * T * T, out_inner_stride=T * T, ldc=T)
    a_q = [act_seg(qk, d, c_total=2 * C)]
    if T <= 256:
        rs = prog.tensor(name + ".rowsum", B * heads * T, "f32")
        prog.gemm(a_q, view(qk, C), B * T, 2 * C, 1, 1, T, T, out_bf16=pm, softmax=1, softmax_scale=scale,
                  rowsum_out=rs, **s_args)
    else:
        rs = None
        logits = prog.tensor(name + ".s", B * heads * T * T, "f32")
        prog.gemm(a_q, view(qk, C), B * T, 2 * C, 1, 1, T, T, out_f32=logits, alpha=scale, **s_args)
        prog.softmax_rows(logits, pm, B * heads * T, T)
    prog.gemm([act_seg(pm, T)], vt, B * C, T, 1, 1, T, d, batch=B * heads, inner=heads, a_batch_rows=heads * T,
              a_inner_rows=T, b_batch_rows=C, b_inner_rows=d, out_batch_stride=T * C, out_inner_stride=d, rowscale=rs,
              out_bf16=o, ldc=C)
    if rec is not None:      # what the data-gradient of this block reads (lower_attention_bwd)
        rec.update(qk=qk, vt=vt, pm=pm, rs=rs)
    return o


def pack_dgrad3x3(w):
    """Conv2d weight [Cout, Cin, 3, 3] -> the B operand of its data-gradient GEMM (the same implicit 3x3 GEMM with the
    taps flipped and in / out swapped): [Cin, 9*Cout] with K index = (ky*3+kx)*Cout + co reading w[co, ci, 2-ky, 2-kx]."""
    co, ci = w.shape[0], w.shape[1]
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co).contiguous()


def transposed(prog, name, src, rows, cols, ld_in, in_batch_stride, batch):
    """bf16 [batch][rows][cols] (row pitch ld_in) -> new tensor [batch][cols][rows]."""
    out = prog.tensor(name, batch * rows * cols, "bf16")
    prog.transpose(src, out, rows, cols, ld_in, rows, batch, in_batch_stride, rows * cols)
    return out


def lower_attention_bwd(prog, name, rec, go, B, T, C, heads, scale):
    """Data-gradient of `lower_attention`: go = dL/d(o) bf16 [B*T, C] -> dq | dk | dv bf16 [B*T, 3C] (head-major columns
    inside each third, as the forward q | k | v). Per (sample, head): dV = P^T dO, dP = dO V^T, dS = P (dP - rowsum(dP P)),
    dQ = scale dS K, dK = scale dS^T Q -- tcgen05 GEMMs over bf16 transposes + the row-wise `softmax_bwd` kernel
    (`attn_small_bwd` for T <= 64)."""
    from .program import view
    d = C // heads
    dqkv = prog.tensor(name + ".dqkv", B * T * 3 * C, "bf16")
    if T <= 64:
        prog.attn_small_bwd(rec["qkv"], go, dqkv, B, T, heads, d, scale)
        return dqkv
    qk, vt, pm, rs = rec["qk"], rec["vt"], rec["pm"], rec["rs"]
    v = transposed(prog, name + ".v", vt, C, T, T, C * T, B)                          # [B][T][C]
    qT = transposed(prog, name + ".qT", view(qk, 0), T, C, 2 * C, T * 2 * C, B)       # [B][C][T]
    kT = transposed(prog, name + ".kT", view(qk, C), T, C, 2 * C, T * 2 * C, B)
    goT = transposed(prog, name + ".goT", go, T, C, C, T * C, B)
    BH = B * heads
    # dP[b,h] = dO[b,:,h] . V[b,:,h]^T   (fp32 [B][heads][T][T])
    dp = prog.tensor(name + ".dp", BH * T * T, "f32")
    prog.gemm([act_seg(go, d, c_total=C)], v, B * T, C, 1, 1, T, T, batch=BH, inner=heads, a_batch_rows=T, a_inner_k=d,
              b_batch_rows=T, b_inner_k=d, w_cols=C, out_batch_stride=heads * T * T, out_inner_stride=T * T, out_f32=dp,
              ldc=T)
    ds = prog.tensor(name + ".ds", BH * T * T, "bf16")
    pn = prog.tensor(name + ".pn", BH * T * T, "bf16")
    prog.softmax_bwd(pm, rs, dp, ds, pn, BH * T, T)            # rs None: pm is the normalised softmax already
    dsT = transposed(prog, name + ".dsT", ds, T, T, T, T * T, BH)
    pnT = transposed(prog, name + ".pnT", pn, T, T, T, T * T, BH)
    # [T, d] results into the head's columns of the q / k / v third of dqkv
    bat = dict(batch=BH, inner=heads, a_batch_rows=heads * T, a_inner_rows=T, b_batch_rows=C, b_inner_rows=d,
               out_batch_stride=T * 3 * C, out_inner_stride=d, ldc=3 * C)
    prog.gemm([act_seg(ds, T)], kT, B * C, T, 1, 1, T, d, alpha=scale, out_bf16=view(dqkv, 0), **bat)
    prog.gemm([act_seg(dsT, T)], qT, B * C, T, 1, 1, T, d, alpha=scale, out_bf16=view(dqkv, C), **bat)
    prog.gemm([act_seg(pnT, T)], goT, B * C, T, 1, 1, T, d, out_bf16=view(dqkv, 2 * C), **bat)
    return dqkv
