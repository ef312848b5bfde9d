"""CPU interpreter of an engine `Program` -- test infrastructure.

Replays the op list the lowering emitted with plain torch ops on flat host buffers, following the
semantics of the CUDA kernels (dp_gemm.cu epilogue order, dp_elem.cu GroupNorm-from-partials, ...).
With emulate_bf16=False it must reproduce the oracle to fp32 round-off, which proves the lowering
(topology, operand offsets, weight packing, statistics plumbing) without a GPU; with emulate_bf16=True
it predicts the tensor-core path's rounding (bf16 operands, fp32 accumulation).
"""
import math

import torch
import torch.nn.functional as F


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


class Interp:
    def __init__(self, prog, emulate_bf16=False):
        self.prog = prog
        self.bf = emulate_bf16
        self.mem = {}
        for t in prog.tensors:
            if t.init is not None:
                v = t.init.detach().float().reshape(-1).clone()
                if t.dtype == "bf16" and self.bf:
                    v = _bf16(v)
                self.mem[t.index] = v
            else:
                self.mem[t.index] = torch.zeros(t.numel)
        self.out = None

    # -- helpers ----------------------------------------------------------------------------------
    def flat(self, v):
        return self.mem[v.tensor.index], v.offset

    def rd(self, v, shape, strides=None):
        buf, off = self.flat(v)
        if strides is None:
            n = 1
            for s in shape:
                n *= s
            return buf[off:off + n].reshape(shape)
        return torch.as_strided(buf, shape, strides, off)

    def wr(self, v, value, shape, strides):
        buf, off = self.flat(v)
        val = _bf16(value) if (self.bf and v.tensor.dtype == "bf16") else value
        torch.as_strided(buf, shape, strides, off).copy_(val)

    # -- ops --------------------------------------------------------------------------------------
    def run(self, x_nchw, cond):
        self.x = x_nchw.float()
        self.cond = cond.float()
        for op in self.prog.ops:
            getattr(self, "op_" + op.kind)(**op.args)
        return self.out

    def op_embed(self, out, B, dim, cos_first, half_minus_1):
        half = dim // 2
        coef = -math.log(10000.0) / (half - 1 if half_minus_1 else half)
        freq = torch.exp(torch.arange(half, dtype=torch.float32) * coef)
        arg = self.cond[:, None] * freq[None]
        e = torch.cat([torch.cos(arg), torch.sin(arg)] if cos_first else [torch.sin(arg), torch.cos(arg)], 1)
        self.wr(out, e, (B, dim), (dim, 1))

    def _a_matrix(self, seg, B, H, W, b, a_batch_rows, k_off=0):
        if H == 1 and seg.taps == 1:  # plain GEMM rows
            buf, off = self.flat(seg.act)
            return torch.as_strided(buf, (W, seg.C), (seg.c_total, 1), off + b * a_batch_rows * seg.c_total + k_off)
        s = seg.stride
        Hin, Win = H * s, W * s
        x = self.rd(seg.act, (B, Hin, Win, seg.C), (Hin * Win * seg.c_total, Win * seg.c_total, seg.c_total, 1))
        cols = []
        k = 3 if seg.taps == 9 else 1
        for ky in range(k):
            for kx in range(k):
                dy, dx = ky - seg.pad, kx - seg.pad
                patch = torch.zeros(B, H, W, seg.C)
                for hh in range(H):
                    yi = hh * s + dy
                    if yi < 0 or yi >= Hin:
                        continue
                    xs = [w * s + dx for w in range(W)]
                    valid = [i for i, xi in enumerate(xs) if 0 <= xi < Win]
                    if not valid:
                        continue
                    w0, w1 = valid[0], valid[-1] + 1
                    patch[:, hh, w0:w1] = x[:, yi, xs[w0]:xs[w1 - 1] + 1:s]
                cols.append(patch.reshape(B * H * W, seg.C))
        return torch.cat(cols, 1)

    def op_gemm(self, a, w, w_rows, w_pitch, B, H, W, N, batch, a_batch_rows, b_batch_rows, out_batch_stride, inner,
                a_inner_k, a_inner_rows, b_inner_k, b_inner_rows, out_inner_stride, w_cols, bias,
                bias_along_m, rowvec, rowvec_ld, rowvec_rows_per_sample, rowscale, resid, alpha, silu, out_f32,
                out_bf16, ldc, stats, softmax, softmax_scale, rowsum_out, gn_out=None, gn_gamma=None, gn_beta=None,
                gn_groups=0, gn_eps=0.0, gn_silu=0):
        M = B * H * W
        ktot = sum(s.taps * s.C for s in a)
        for bi in range(batch):
            b, hd = bi // inner, bi % inner
            obs = b * out_batch_stride + hd * out_inner_stride
            A = torch.cat([self._a_matrix(s, B, H, W, b, a_batch_rows, hd * a_inner_k + hd * a_inner_rows * s.c_total)
                           for s in a], 1)
            wbuf, woff = self.flat(w)
            row0 = b * b_batch_rows + hd * b_inner_rows
            nvalid = max(0, min(N, w_rows - row0))          # rows past the matrix read as zeros (TMA OOB fill)
            Wt = torch.zeros(N, ktot)
            Wt[:nvalid] = torch.as_strided(wbuf, (nvalid, ktot), (w_pitch, 1), woff + row0 * w_pitch + hd * b_inner_k)
            D = A @ Wt.t()
            if softmax:
                mx = D.max(dim=1, keepdim=True).values
                Pm = torch.exp((D - mx) * softmax_scale)
                if self.bf:
                    Pm = _bf16(Pm)
                buf, off = self.flat(out_bf16)
                torch.as_strided(buf, (M, N), (ldc, 1), off + obs).copy_(Pm)
                rbuf, roff = self.flat(rowsum_out)
                rbuf[roff + bi * M: roff + (bi + 1) * M] = Pm.sum(1)
                continue
            if rowscale is not None:
                rbuf, roff = self.flat(rowscale)
                D = D / rbuf[roff + bi * M: roff + (bi + 1) * M][:, None]
            if bias is not None:
                bb, boff = self.flat(bias)
                D = D + (bb[boff:boff + M][:, None] if bias_along_m else bb[boff:boff + N][None, :])
            if rowvec is not None:
                rv, rvo = self.flat(rowvec)
                nsamp = M // rowvec_rows_per_sample
                tab = torch.as_strided(rv, (nsamp, N), (rowvec_ld, 1), rvo)
                D = D + tab.repeat_interleave(rowvec_rows_per_sample, 0)
            if silu:
                D = F.silu(D)
            if resid is not None:
                rb, ro = self.flat(resid)
                D = D + torch.as_strided(rb, (M, N), (ldc, 1), ro + obs)
            D = D * alpha
            if gn_out is not None:        # fused GroupNorm(+SiLU) epilogue: statistics of the fp32 result per sample
                assert batch == 1 and (out_f32 is not None or (out_bf16 is None and stats is None and resid is None))
                x = D.reshape(B, H * W, gn_groups, N // gn_groups)
                mean = x.double().mean((1, 3), keepdim=True)
                var = (x.double() * x.double()).mean((1, 3), keepdim=True) - mean * mean
                y = ((x - mean.float()) * (1.0 / torch.sqrt(var.clamp_min(0) + gn_eps)).float()).reshape(M, N)
                y = y * self.rd(gn_gamma, (N,))[None] + self.rd(gn_beta, (N,))[None]
                if gn_silu:
                    y = F.silu(y)
                buf, off = self.flat(gn_out)
                torch.as_strided(buf, (M, N), (ldc, 1), off + obs).copy_(_bf16(y) if self.bf else y)
            for dst in (out_f32, out_bf16):
                if dst is not None:
                    buf, off = self.flat(dst)
                    val = _bf16(D) if (self.bf and dst.tensor.dtype == "bf16") else D
                    torch.as_strided(buf, (M, N), (ldc, 1), off + obs).copy_(val)
            if stats is not None:
                seg = min(H * W, 128)
                nseg = M // seg
                Ds = D.reshape(nseg, seg, N)
                st = torch.stack([Ds.sum(1), (Ds * Ds).sum(1)], -1)  # [nseg, N, 2]
                sb, so = self.flat(stats)
                sb[so: so + st.numel()] = st.reshape(-1)

    def op_gn_apply(self, src0, stats0, C0, P0, src1, stats1, C1, P1, gamma, beta, film, film_ld, B, H, W, groups,
                    eps, silu, resample, out_bf16, raw_bf16, raw_f32):
        C = C0 + C1
        HW = H * W
        if stats0 is None:   # identity: cast / resample only
            y = self.rd(src0, (B, H, W, C0))
            if resample == 1:
                y = y.repeat_interleave(2, 1).repeat_interleave(2, 2)
            elif resample == 2:
                y = y.reshape(B, H // 2, 2, W // 2, 2, C).mean((2, 4))
            Ho, Wo = y.shape[1], y.shape[2]
            self.wr(out_bf16, y, (B, Ho, Wo, C), (Ho * Wo * C, Wo * C, C, 1))
            return
        xs = [self.rd(src0, (B, H, W, C0))]
        sums = [self.rd(stats0, (B, P0, C0, 2)).sum(1)]
        if src1 is not None:
            xs.append(self.rd(src1, (B, H, W, C1)))
            sums.append(self.rd(stats1, (B, P1, C1, 2)).sum(1))
        x = torch.cat(xs, -1)
        cs = torch.cat(sums, 1).double()                      # [B, C, 2]
        cpg = C // groups
        gsum = cs.reshape(B, groups, cpg, 2).sum(2)
        n = cpg * HW
        mean = gsum[..., 0] / n
        var = (gsum[..., 1] / n - mean * mean).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var + eps)
        g, bt = self.rd(gamma, (C,)), self.rd(beta, (C,))
        sc = g[None] * rstd.float().repeat_interleave(cpg, 1)
        sh = bt[None] - mean.float().repeat_interleave(cpg, 1) * sc
        if film is not None:
            fb, fo = self.flat(film)
            f = torch.as_strided(fb, (B, 2 * C), (film_ld, 1), fo)
            fs = 1.0 + f[:, :C]
            sc, sh = sc * fs, sh * fs + f[:, C:]
        y = x * sc[:, None, None, :] + sh[:, None, None, :]
        if silu:
            y = F.silu(y)
        raw = x
        if resample == 1:
            y = y.repeat_interleave(2, 1).repeat_interleave(2, 2)
            raw = raw.repeat_interleave(2, 1).repeat_interleave(2, 2)
        elif resample == 2:
            y = y.reshape(B, H // 2, 2, W // 2, 2, C).mean((2, 4))
            raw = raw.reshape(B, H // 2, 2, W // 2, 2, C).mean((2, 4))
        Ho, Wo = y.shape[1], y.shape[2]
        st = (Ho * Wo * C, Wo * C, C, 1)
        self.wr(out_bf16, y, (B, Ho, Wo, C), st)
        if raw_bf16 is not None:
            self.wr(raw_bf16, raw, (B, Ho, Wo, C), st)
        if raw_f32 is not None:
            self.wr(raw_f32, raw, (B, Ho, Wo, C), st)

    def op_conv_in(self, w, bias, out, stats, B, H, W, Cout):
        wt = self.rd(w, (3, 3, 3, Cout)).permute(3, 2, 0, 1).contiguous()   # [(ky,kx,ci), co] -> [co, ci, ky, kx]
        y = F.conv2d(self.x, wt, self.rd(bias, (Cout,)), padding=1).permute(0, 2, 3, 1).contiguous()
        self.wr(out, y, (B, H, W, Cout), (H * W * Cout, W * Cout, Cout, 1))
        if stats is not None:
            HW = H * W
            P = (HW + 127) // 128
            yy = y.reshape(B, HW, Cout)
            st = torch.zeros(B, P, Cout, 2)
            for p in range(P):
                blk = yy[:, p * 128:(p + 1) * 128]
                st[:, p, :, 0] = blk.sum(1)
                st[:, p, :, 1] = (blk * blk).sum(1)
            sb, so = self.flat(stats)
            sb[so:so + st.numel()] = st.reshape(-1)

    def op_pad_in(self, out, B, H, W, Cpad):
        x = torch.zeros(B, H, W, Cpad)
        x[..., :3] = self.x.permute(0, 2, 3, 1)
        self.wr(out, x, (B, H, W, Cpad), (H * W * Cpad, W * Cpad, Cpad, 1))

    def op_update(self, eps, ld, B, H, W, Cout):
        e = self.rd(eps, (B, H, W, Cout), (H * W * ld, W * ld, ld, 1))
        self.out = e.permute(0, 3, 1, 2).contiguous()

    def op_softmax_rows(self, src, out, rows, T):
        x = self.rd(src, (rows, T))
        self.wr(out, torch.softmax(x, -1), (rows, T), (T, 1))

    def op_attn_block(self, hn, w, bias, resid, out_f32, stats, B, T, C, scale, alpha):
        """dp_attn.cu: q, k, v^T, P and o are rounded to bf16 where the kernel turns accumulators into operand tiles."""
        r = _bf16 if self.bf else (lambda t: t)
        h = self.rd(hn, (B, T, C))
        W = self.rd(w, (4, C, C))
        b = self.rd(bias, (4, C))
        x = self.rd(resid, (B, T, C))
        q = r((h @ W[0].t() + b[0]) * scale)
        k = r(h @ W[1].t() + b[1])
        v = r(h @ W[2].t() + b[2])
        s = q @ k.transpose(1, 2)
        p = r(torch.exp(s - s.amax(-1, keepdim=True)))
        o = r((p @ v) / p.sum(-1, keepdim=True))
        y = (o @ W[3].t() + b[3] + x) * alpha
        self.wr(out_f32, y, (B, T, C), (T * C, C, 1))
        if stats is not None:
            t = y.reshape(B * T // 128, 128, C)
            self.wr(stats, torch.stack([t.sum(1), (t * t).sum(1)], -1), (B * T // 128, C, 2), (C * 2, 2, 1))

    def op_attn_small(self, qkv, out, B, T, heads, d, scale):
        x = self.rd(qkv, (B, T, 3, heads, d))
        q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]                        # [B, T, heads, d]
        s = torch.einsum("bihd,bjhd->bhij", q, k) * scale
        p = torch.softmax(s, -1)
        o = torch.einsum("bhij,bjhd->bihd", p, v).reshape(B, T, heads * d)
        self.wr(out, o, (B, T, heads * d), (T * heads * d, heads * d, 1))

    # -- data-gradient ops (dp_bwd.cu) --------------------------------------------------------------
    def run_vjp(self, x_nchw, cond, g_nchw):
        self.g_in = g_nchw.float()
        return self.run(x_nchw, cond)

    def op_grad_in(self, out, B, H, W, C, Cpad):
        g = torch.zeros(B, H, W, Cpad)
        g[..., :C] = self.g_in.permute(0, 2, 3, 1)
        self.wr(out, g, (B, H, W, Cpad), (H * W * Cpad, W * Cpad, Cpad, 1))

    @staticmethod
    def _resample_T(t, resample):
        """transpose of the forward resample; t: [B, Ho, Wo, C] at the forward output resolution."""
        if resample == 1:      # forward nearest x2 -> sum of the 2x2 children
            B, Ho, Wo, C = t.shape
            return t.reshape(B, Ho // 2, 2, Wo // 2, 2, C).sum((2, 4))
        if resample == 2:      # forward 2x2 mean -> a quarter of the parent
            return t.repeat_interleave(2, 1).repeat_interleave(2, 2) * 0.25
        return t

    def op_gn_bwd(self, src0, stats0, C0, P0, src1, stats1, C1, P1, gamma, beta, B, H, W, groups, eps, silu, resample,
                  g, add0, add0_scale, add1, d0_f32, d0_bf16, d1_f32, film=None, film_ld=0):
        C = C0 + C1
        HW = H * W
        Ho, Wo = (2 * H, 2 * W) if resample == 1 else ((H // 2, W // 2) if resample == 2 else (H, W))
        xs = [self.rd(src0, (B, H, W, C0))]
        sums = [self.rd(stats0, (B, P0, C0, 2)).sum(1)]
        if src1 is not None:
            xs.append(self.rd(src1, (B, H, W, C1)))
            sums.append(self.rd(stats1, (B, P1, C1, 2)).sum(1))
        x = torch.cat(xs, -1)
        cs = torch.cat(sums, 1).double()
        cpg = C // groups
        gsum = cs.reshape(B, groups, cpg, 2).sum(2)
        n = cpg * HW
        mean = gsum[..., 0] / n
        var = (gsum[..., 1] / n - mean * mean).clamp_min(0)
        rstd = (1.0 / torch.sqrt(var + eps)).float().repeat_interleave(cpg, 1)[:, None, None, :]
        mean = mean.float().repeat_interleave(cpg, 1)[:, None, None, :]
        xhat = (x - mean) * rstd
        gam = self.rd(gamma, (C,))[None, None, None, :]
        bet = self.rd(beta, (C,))[None, None, None, :]
        if film is not None:   # scale-shift norm: per-sample (1 + scale), shift fold into gamma, beta
            fb, fo = self.flat(film)
            f = torch.as_strided(fb, (B, 2 * C), (film_ld, 1), fo)
            sc = 1 + f[:, None, None, :C]
            bet = bet * sc + f[:, None, None, C:]
            gam = gam * sc
        gy = self._resample_T(self.rd(g, (B, Ho, Wo, C)), resample)
        if silu:
            u = xhat * gam + bet
            s = torch.sigmoid(u)
            gy = gy * (s * (1 + u * (1 - s)))
        gx = gy * gam
        gxg = gx.reshape(B, HW, groups, cpg)
        xhg = xhat.reshape(B, HW, groups, cpg)
        m1 = gxg.mean((1, 3)).repeat_interleave(cpg, 1)[:, None, None, :]
        m2 = (gxg * xhg).mean((1, 3)).repeat_interleave(cpg, 1)[:, None, None, :]
        d = rstd * (gx - m1 - xhat * m2)
        if add0 is not None:
            d = d + add0_scale * self._resample_T(self.rd(add0, (B, Ho, Wo, C)), resample)
        d0 = d[..., :C0]
        if add1 is not None:
            d0 = d0 + self.rd(add1, (B, H, W, C0))
        st0 = (HW * C0, W * C0, C0, 1)
        if d0_f32 is not None:
            self.wr(d0_f32, d0, (B, H, W, C0), st0)
        if d0_bf16 is not None:
            self.wr(d0_bf16, d0, (B, H, W, C0), st0)
        if C1:
            self.wr(d1_f32, d[..., C0:], (B, H, W, C1), (HW * C1, W * C1, C1, 1))

    def op_softmax_bwd(self, pnum, rowsum, dp, ds, pn, rows, T):
        P = self.rd(pnum, (rows, T))
        if rowsum is not None:
            P = P / self.rd(rowsum, (rows,))[:, None]
        dP = self.rd(dp, (rows, T))
        dS = P * (dP - (dP * P).sum(-1, keepdim=True))
        self.wr(ds, dS, (rows, T), (T, 1))
        self.wr(pn, P, (rows, T), (T, 1))

    def op_transpose(self, src, out, rows, cols, ld_in, ld_out, batch, in_batch_stride, out_batch_stride):
        buf, off = self.flat(src)
        x = torch.as_strided(buf, (batch, rows, cols), (in_batch_stride, ld_in, 1), off)
        self.wr(out, x.transpose(1, 2), (batch, cols, rows), (out_batch_stride, ld_out, 1))

    def op_attn_small_bwd(self, qkv, go, out, B, T, heads, d, scale):
        x = self.rd(qkv, (B, T, 3, heads, d))
        q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
        g = self.rd(go, (B, T, heads, d))
        p = torch.softmax(torch.einsum("bihd,bjhd->bhij", q, k) * scale, -1)
        dv = torch.einsum("bhij,bihd->bjhd", p, g)
        dp = torch.einsum("bihd,bjhd->bhij", g, v)
        ds = p * (dp - (dp * p).sum(-1, keepdim=True))
        dq = torch.einsum("bhij,bjhd->bihd", ds, k) * scale
        dk = torch.einsum("bhij,bihd->bjhd", ds, q) * scale
        o = torch.stack([dq, dk, dv], 2).reshape(B, T, 3 * heads * d)
        self.wr(out, o, (B, T, 3 * heads * d), (T * 3 * heads * d, 3 * heads * d, 1))
