"""Abstract program: the score-model UNet lowered to the engine's fused ops.

The lowering (diffpure_b200/lowering_*.py) emits device tensors and ops into a `Program`; the engine
backend (diffpure_b200/engine.py) materialises it through the C ABI (include/diffpure_b200.h). The op
records mirror the C descriptor structs one to one, so the same program can be replayed by the CPU
interpreter used in the host-logic tests.
"""
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple


@dataclass
class Tensor:
    """A device tensor. `init` (torch tensor) marks constants (weights, tables) uploaded once."""
    name: str
    numel: int
    dtype: str                      # 'f32' | 'bf16'
    init: Any = None
    index: int = -1

    @property
    def nbytes(self):
        return self.numel * (4 if self.dtype == "f32" else 2)


@dataclass
class View:
    """`tensor` seen from an element offset (pointer arithmetic on the device)."""
    tensor: Tensor
    offset: int = 0


def view(t, offset=0):
    if t is None:
        return None
    if isinstance(t, View):
        return View(t.tensor, t.offset + offset)
    return View(t, offset)


@dataclass
class ASeg:
    act: View          # bf16 NHWC [B,Hin,Win,c_total] (plain GEMM: [rows, c_total])
    C: int
    c_total: int
    taps: int = 1
    stride: int = 1
    pad: int = 0


@dataclass
class Op:
    kind: str
    args: Dict[str, Any] = field(default_factory=dict)


class Program:
    def __init__(self, batch, height, width):
        self.B, self.H, self.W = batch, height, width
        self.tensors: List[Tensor] = []
        self.ops: List[Op] = []
        self.meta: Dict[str, Any] = {}

    # -- tensors --------------------------------------------------------------------------------
    def tensor(self, name, numel, dtype, init=None):
        t = Tensor(name, int(numel), dtype, init, len(self.tensors))
        self.tensors.append(t)
        return t

    def const_f32(self, name, value):
        return self.tensor(name, value.numel(), "f32", value.detach().float().contiguous().reshape(-1))

    def const_bf16(self, name, value):
        return self.tensor(name, value.numel(), "bf16", value.detach().float().contiguous().reshape(-1))

    # -- ops ------------------------------------------------------------------------------------
    def add(self, kind, **args):
        self.ops.append(Op(kind, args))

    def embed(self, out, B, dim, cos_first, half_minus_1):
        self.add("embed", out=view(out), B=B, dim=dim, cos_first=cos_first, half_minus_1=half_minus_1)

    def gemm(self, a: List[ASeg], w, w_rows, w_pitch, B, H, W, N, *, batch=1, a_batch_rows=0, b_batch_rows=0,
             out_batch_stride=0, inner=1, a_inner_k=0, a_inner_rows=0, b_inner_k=0, b_inner_rows=0, out_inner_stride=0, w_cols=0,
             bias=None, bias_along_m=0, rowvec=None, rowvec_ld=0, rowvec_rows_per_sample=1,
             rowscale=None, resid=None, alpha=1.0, silu=0, out_f32=None, out_bf16=None, ldc=None, stats=None,
             softmax=0, softmax_scale=1.0, rowsum_out=None, gn_out=None, gn_gamma=None, gn_beta=None, gn_groups=0,
             gn_eps=0.0, gn_silu=0):
        """gn_out: bf16 tensor receiving act(GroupNorm(result)) -- the fused GroupNorm epilogue; the raw result is then not
        written at all (out_f32 / out_bf16 / stats must be None)."""
        self.add("gemm", a=a, w=view(w), w_rows=w_rows, w_pitch=w_pitch, B=B, H=H, W=W, N=N, batch=batch,
                 a_batch_rows=a_batch_rows, b_batch_rows=b_batch_rows, out_batch_stride=out_batch_stride,
                 inner=inner, a_inner_k=a_inner_k, a_inner_rows=a_inner_rows, b_inner_k=b_inner_k, b_inner_rows=b_inner_rows,
                 out_inner_stride=out_inner_stride, w_cols=w_cols, bias=view(bias), bias_along_m=bias_along_m, rowvec=view(rowvec), rowvec_ld=rowvec_ld,
                 rowvec_rows_per_sample=rowvec_rows_per_sample, rowscale=view(rowscale), resid=view(resid),
                 alpha=float(alpha), silu=silu, out_f32=view(out_f32), out_bf16=view(out_bf16),
                 ldc=N if ldc is None else ldc, stats=view(stats), softmax=softmax,
                 softmax_scale=float(softmax_scale), rowsum_out=view(rowsum_out), gn_out=view(gn_out),
                 gn_gamma=view(gn_gamma), gn_beta=view(gn_beta), gn_groups=gn_groups, gn_eps=float(gn_eps),
                 gn_silu=gn_silu)

    def cast(self, *, src, C, B, H, W, out_bf16, resample=0):
        """Identity gn_apply: bf16 copy (optionally resampled) of a raw fp32 stream tensor."""
        self.add("gn_apply", src0=view(src), stats0=None, C0=C, P0=0, src1=None, stats1=None, C1=0, P1=0, gamma=None,
                 beta=None, film=None, film_ld=0, B=B, H=H, W=W, groups=1, eps=0.0, silu=0, resample=resample,
                 out_bf16=view(out_bf16), raw_bf16=None, raw_f32=None)

    def softmax_rows(self, src, out, rows, T):
        self.add("softmax_rows", src=view(src), out=view(out), rows=rows, T=T)

    def gn_apply(self, *, src0, stats0, C0, P0, gamma, beta, B, H, W, groups, eps, silu, out_bf16, src1=None,
                 stats1=None, C1=0, P1=0, film=None, film_ld=0, resample=0, raw_bf16=None, raw_f32=None):
        self.add("gn_apply", src0=view(src0), stats0=view(stats0), C0=C0, P0=P0, src1=view(src1),
                 stats1=view(stats1), C1=C1, P1=P1, gamma=view(gamma), beta=view(beta), film=view(film),
                 film_ld=film_ld, B=B, H=H, W=W, groups=groups, eps=float(eps), silu=silu, resample=resample,
                 out_bf16=view(out_bf16), raw_bf16=view(raw_bf16), raw_f32=view(raw_f32))

    def conv_in(self, w, bias, out, stats, B, H, W, Cout):
        self.add("conv_in", w=view(w), bias=view(bias), out=view(out), stats=view(stats), B=B, H=H, W=W, Cout=Cout)

    def conv_in_gemm(self, name, w_oc33, bias, out, stats, B, H, W, Cout):
        """The 3 -> C input conv on the tensor cores: the state is cast to bf16 and zero-padded to 64 channels (`pad_in`),
        the weights to K = 9 * 64 (27 non-zero columns); bias, fp32 output and the GroupNorm partial statistics come from the
        GEMM epilogue (no separate statistics kernel)."""
        import torch
        from .lowering_common import act_seg, pack_conv3x3
        xin = self.tensor(name + ".x64", B * H * W * 64, "bf16")
        self.add("pad_in", out=view(xin), B=B, H=H, W=W, Cpad=64)
        w64 = torch.zeros(Cout, 64, 3, 3)
        w64[:, :w_oc33.shape[1]] = w_oc33.detach().float().cpu()
        self.gemm([act_seg(xin, 64, taps=9)], self.const_bf16(name + ".w", pack_conv3x3(w64)), Cout, 9 * 64, B, H, W, Cout,
                  bias=self.const_f32(name + ".b", bias), out_f32=out, stats=stats)

    def update(self, eps, ld, B, H, W, Cout):
        self.add("update", eps=view(eps), ld=ld, B=B, H=H, W=W, Cout=Cout)

    def conv_out_gemm(self, name, act, w_oc33, bias, B, H, W, C, Cout):
        """The C->3|6 output conv on the tensor cores (N padded to 8 columns; rows >= Cout of the weight matrix are
        read as zeros through TMA out-of-bounds fill) followed by the fused per-step update."""
        import torch
        from .lowering_common import act_seg, pack_conv3x3
        eps = self.tensor(name + ".eps", B * H * W * 8, "f32")
        b8 = torch.cat([bias.detach().float().cpu(), torch.zeros(8 - Cout)])
        self.gemm([act_seg(act, C, taps=9)], self.const_bf16(name + ".w", pack_conv3x3(w_oc33)), Cout, 9 * C, B, H, W, 8,
                  bias=self.const_f32(name + ".b", b8), out_f32=eps, ldc=8)
        self.update(eps, 8, B, H, W, Cout)

    # -- data-gradient ops (input gradient of the network; diffpure_b200/csrc/dp_bwd.cu) -------------------
    def gn_bwd(self, *, src0, stats0, C0, P0, gamma, beta, B, H, W, groups, eps, silu, g, src1=None, stats1=None, C1=0,
               P1=0, resample=0, add0=None, add0_scale=1.0, add1=None, d0_f32=None, d0_bf16=None, d1_f32=None,
               film=None, film_ld=0):
        self.add("gn_bwd", film=view(film), film_ld=film_ld, src0=view(src0), stats0=view(stats0), C0=C0, P0=P0, src1=view(src1), stats1=view(stats1),
                 C1=C1, P1=P1, gamma=view(gamma), beta=view(beta), B=B, H=H, W=W, groups=groups, eps=float(eps),
                 silu=silu, resample=resample, g=view(g), add0=view(add0), add0_scale=float(add0_scale),
                 add1=view(add1), d0_f32=view(d0_f32), d0_bf16=view(d0_bf16), d1_f32=view(d1_f32))

    def softmax_bwd(self, pnum, rowsum, dp, ds, pn, rows, T):
        self.add("softmax_bwd", pnum=view(pnum), rowsum=view(rowsum), dp=view(dp), ds=view(ds), pn=view(pn), rows=rows,
                 T=T)

    def transpose(self, src, out, rows, cols, ld_in, ld_out, batch, in_batch_stride, out_batch_stride):
        self.add("transpose", src=view(src), out=view(out), rows=rows, cols=cols, ld_in=ld_in, ld_out=ld_out,
                 batch=batch, in_batch_stride=in_batch_stride, out_batch_stride=out_batch_stride)

    def attn_small_bwd(self, qkv, go, out, B, T, heads, d, scale):
        self.add("attn_small_bwd", qkv=view(qkv), go=view(go), out=view(out), B=B, T=T, heads=heads, d=d,
                 scale=float(scale))

    def grad_in(self, out, B, H, W, C, Cpad):
        self.add("grad_in", out=view(out), B=B, H=H, W=W, C=C, Cpad=Cpad)

    def attn_block(self, hn, w, bias, resid, out_f32, stats, B, T, C, scale, alpha):
        """AttnBlockpp behind its GroupNorm as one kernel (layerspp.py:75-91): w = Wq | Wk | Wv | W3 as [4C, C] ([out, in]),
        bias = bq | bk | bv | b3; out = (resid + NIN_3(softmax(q k^T scale) v)) * alpha (+ partial statistics of out)."""
        self.add("attn_block", hn=view(hn), w=view(w), bias=view(bias), resid=view(resid), out_f32=view(out_f32),
                 stats=view(stats), B=B, T=T, C=C, scale=float(scale), alpha=float(alpha))

    def attn_small(self, qkv, out, B, T, heads, d, scale):
        self.add("attn_small", qkv=view(qkv), out=view(out), B=B, T=T, heads=heads, d=d, scale=float(scale))

    # -- analysis -------------------------------------------------------------------------------
    def views_of(self, op):
        """All (View, is_written) operands of an op, for liveness analysis."""
        out = []
        for k, v in op.args.items():
            if isinstance(v, View):
                out.append(v)
            elif k == "a":
                out.extend(seg.act for seg in v)
        return out

    def last_use(self):
        last = {}
        for i, op in enumerate(self.ops):
            for v in self.views_of(op):
                last[v.tensor.index] = i
        return last

    def first_use(self):
        first = {}
        for i, op in enumerate(self.ops):
            for v in self.views_of(op):
                first.setdefault(v.tensor.index, i)
        return first


def stats_rows(B, HW):
    """Rows of a [rows][C][2] GroupNorm partial-statistics tensor written by a GEMM epilogue or conv_in."""
    if HW >= 128:
        return B * (HW // 128), HW // 128
    ipt = 128 // HW
    return ((B + ipt - 1) // ipt) * ipt, 1
