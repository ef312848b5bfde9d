"""Lowering of the CelebA-HQ DDPM UNet (SDEdit checkpoint family) to the engine program.

Mirrors ddpm/unet_ddpm.py:200-345: ResnetBlock (L85-142, additive temb, no output rescale), AttnBlock (L145-197,
single head of width C, 1x1-conv projections), conv resampling (Downsample = pad (0,1,0,1) + 3x3 stride 2, L63-82;
Upsample = nearest x2 + 3x3, L44-60), GroupNorm(32, eps 1e-6), [sin | cos] timestep embedding (L14-32).
State-dict names are the reference's.
"""
from types import SimpleNamespace

import torch

from .lowering_common import Act, act_seg, lower_attention, new_act, pack_conv1x1, pack_conv3x3, pack_conv_in, \
    pad_rows
from .program import Program, view

EPS = 1e-6


def celeba_cfg():
    return SimpleNamespace(image_size=256, ch=128, out_ch=3, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2,
                           attn_resolutions=(16,), in_channels=3)


def cfg_from_reference(config):
    m = config.model
    assert m.resamp_with_conv and m.in_channels == 3, "unsupported DDPM UNet variant"
    return SimpleNamespace(image_size=config.data.image_size, ch=m.ch, out_ch=m.out_ch, ch_mult=tuple(m.ch_mult),
                           num_res_blocks=m.num_res_blocks, attn_resolutions=tuple(m.attn_resolutions), in_channels=3)


def block_list(cfg):
    """Every ResnetBlock in execution order: (prefix, cin, cout) -- shared by the temb GEMM and the walk."""
    ch, nres = cfg.ch, len(cfg.ch_mult)
    in_mult = (1,) + tuple(cfg.ch_mult)
    blocks = []
    block_in = None
    for lvl in range(nres):
        block_in, block_out = ch * in_mult[lvl], ch * cfg.ch_mult[lvl]
        for b in range(cfg.num_res_blocks):
            blocks.append((f"down.{lvl}.block.{b}.", block_in, block_out))
            block_in = block_out
    blocks.append(("mid.block_1.", block_in, block_in))
    blocks.append(("mid.block_2.", block_in, block_in))
    for lvl in reversed(range(nres)):
        block_out = ch * cfg.ch_mult[lvl]
        skip_in = ch * cfg.ch_mult[lvl]
        for b in range(cfg.num_res_blocks + 1):
            if b == cfg.num_res_blocks:
                skip_in = ch * in_mult[lvl]
            blocks.append((f"up.{lvl}.block.{b}.", block_in + skip_in, block_out))
            block_in = block_out
    return blocks


def param_shapes(cfg):
    sh = {}
    ch, temb = cfg.ch, cfg.ch * 4
    sh["temb.dense.0.weight"], sh["temb.dense.0.bias"] = (temb, ch), (temb,)
    sh["temb.dense.1.weight"], sh["temb.dense.1.bias"] = (temb, temb), (temb,)
    sh["conv_in.weight"], sh["conv_in.bias"] = (ch, 3, 3, 3), (ch,)
    for p, cin, cout in block_list(cfg):
        sh[p + "norm1.weight"], sh[p + "norm1.bias"] = (cin,), (cin,)
        sh[p + "conv1.weight"], sh[p + "conv1.bias"] = (cout, cin, 3, 3), (cout,)
        sh[p + "temb_proj.weight"], sh[p + "temb_proj.bias"] = (cout, temb), (cout,)
        sh[p + "norm2.weight"], sh[p + "norm2.bias"] = (cout,), (cout,)
        sh[p + "conv2.weight"], sh[p + "conv2.bias"] = (cout, cout, 3, 3), (cout,)
        if cin != cout:
            sh[p + "nin_shortcut.weight"], sh[p + "nin_shortcut.bias"] = (cout, cin, 1, 1), (cout,)
    nres = len(cfg.ch_mult)
    res = cfg.image_size
    in_mult = (1,) + tuple(cfg.ch_mult)

    def attn(p, c):
        sh[p + "norm.weight"], sh[p + "norm.bias"] = (c,), (c,)
        for n in ("q", "k", "v", "proj_out"):
            sh[p + n + ".weight"], sh[p + n + ".bias"] = (c, c, 1, 1), (c,)

    for lvl in range(nres):
        c = ch * cfg.ch_mult[lvl]
        if res in cfg.attn_resolutions:
            for b in range(cfg.num_res_blocks):
                attn(f"down.{lvl}.attn.{b}.", c)
        if lvl != nres - 1:
            sh[f"down.{lvl}.downsample.conv.weight"], sh[f"down.{lvl}.downsample.conv.bias"] = (c, c, 3, 3), (c,)
            res //= 2
    attn("mid.attn_1.", ch * cfg.ch_mult[-1])
    for lvl in reversed(range(nres)):
        c = ch * cfg.ch_mult[lvl]
        if res in cfg.attn_resolutions:
            for b in range(cfg.num_res_blocks + 1):
                attn(f"up.{lvl}.attn.{b}.", c)
        if lvl != 0:
            sh[f"up.{lvl}.upsample.conv.weight"], sh[f"up.{lvl}.upsample.conv.bias"] = (c, c, 3, 3), (c,)
            res *= 2
    sh["norm_out.weight"], sh["norm_out.bias"] = (ch * cfg.ch_mult[0],), (ch * cfg.ch_mult[0],)
    sh["conv_out.weight"], sh["conv_out.bias"] = (cfg.out_ch, ch * cfg.ch_mult[0], 3, 3), (cfg.out_ch,)
    return sh


def lower(cfg, sd, B, h_bf16=True):
    """h_bf16: conv1's output (read only by norm2) is stored in bf16."""
    S = cfg.image_size
    prog = Program(B, S, S)
    ch, temb_dim = cfg.ch, cfg.ch * 4
    nres = len(cfg.ch_mult)

    def P(name):
        return sd[name].detach().float().cpu()

    # ---- timestep embedding MLP + all temb_proj(swish(temb)) in one GEMM ------------------------------
    blocks = block_list(cfg)
    dense_off, off = {}, 0
    for p, cin, cout in blocks:
        dense_off[p] = off
        off += cout
    n_all = (off + 127) // 128 * 128
    w_all = pad_rows(torch.cat([P(p + "temb_proj.weight") for p, _, _ in blocks], 0))
    b_all = torch.cat([P(p + "temb_proj.bias") for p, _, _ in blocks] + [torch.zeros(n_all - off)], 0)
    emb = prog.tensor("temb.emb", B * ch, "bf16")
    prog.embed(emb, B, ch, cos_first=0, half_minus_1=1)
    t1 = prog.tensor("temb.h1", B * temb_dim, "bf16")
    prog.gemm([act_seg(emb, ch)], prog.const_bf16("temb.w0", P("temb.dense.0.weight")), temb_dim, ch, 1, 1, B, temb_dim,
              bias=prog.const_f32("temb.b0", P("temb.dense.0.bias")), silu=1, out_bf16=t1)
    t2 = prog.tensor("temb.h2", B * temb_dim, "bf16")
    prog.gemm([act_seg(t1, temb_dim)], prog.const_bf16("temb.w1", P("temb.dense.1.weight")), temb_dim, temb_dim, 1, 1,
              B, temb_dim, bias=prog.const_f32("temb.b1", P("temb.dense.1.bias")), silu=1, out_bf16=t2)
    temb_all = prog.tensor("temb.all", B * n_all, "f32")
    prog.gemm([act_seg(t2, temb_dim)], prog.const_bf16("temb.wall", w_all), n_all, temb_dim, 1, 1, B, n_all,
              bias=prog.const_f32("temb.ball", b_all), out_f32=temb_all)

    def resblock(p, x0: Act, x1: Act = None):
        """ResnetBlock.forward, unet_ddpm.py:123-142."""
        cin = x0.C + (x1.C if x1 else 0)
        cout = P(p + "conv1.weight").shape[0]
        H, W = x0.H, x0.W
        shortcut = cin != cout
        a0 = prog.tensor(p + "a0", B * H * W * cin, "bf16")
        xb = prog.tensor(p + "xb", B * H * W * cin, "bf16") if shortcut else None
        prog.gn_apply(src0=x0.t, stats0=x0.stats, C0=x0.C, P0=x0.P, src1=x1.t if x1 else None,
                      stats1=x1.stats if x1 else None, C1=x1.C if x1 else 0, P1=x1.P if x1 else 0,
                      gamma=prog.const_f32(p + "n1.w", P(p + "norm1.weight")),
                      beta=prog.const_f32(p + "n1.b", P(p + "norm1.bias")), B=B, H=H, W=W, groups=32, eps=EPS, silu=1,
                      out_bf16=a0, raw_bf16=xb)
        h = new_act(prog, p + "h", B, cout, H, W)
        if h_bf16:
            h.t = prog.tensor(p + "h16", B * H * W * cout, "bf16")
        prog.gemm([act_seg(a0, cin, taps=9)], prog.const_bf16(p + "w1", pack_conv3x3(P(p + "conv1.weight"))), cout,
                  9 * cin, B, H, W, cout, bias=prog.const_f32(p + "b1", P(p + "conv1.bias")),
                  rowvec=view(temb_all, dense_off[p]), rowvec_ld=n_all, rowvec_rows_per_sample=H * W,
                  out_f32=None if h_bf16 else h.t, out_bf16=h.t if h_bf16 else None, stats=h.stats)
        a1 = prog.tensor(p + "a1", B * H * W * cout, "bf16")
        prog.gn_apply(src0=h.t, stats0=h.stats, C0=cout, P0=h.P,
                      gamma=prog.const_f32(p + "n2.w", P(p + "norm2.weight")),
                      beta=prog.const_f32(p + "n2.b", P(p + "norm2.bias")), B=B, H=H, W=W, groups=32, eps=EPS, silu=1,
                      out_bf16=a1)
        out = new_act(prog, p + "out", B, cout, H, W)
        w2 = pack_conv3x3(P(p + "conv2.weight"))
        if shortcut:
            w = torch.cat([w2, pack_conv1x1(P(p + "nin_shortcut.weight"))], dim=1)
            bias = P(p + "conv2.bias") + P(p + "nin_shortcut.bias")
            prog.gemm([act_seg(a1, cout, taps=9), act_seg(xb, cin)], prog.const_bf16(p + "w2", w), cout,
                      9 * cout + cin, B, H, W, cout, bias=prog.const_f32(p + "b2", bias), out_f32=out.t,
                      stats=out.stats)
        else:
            prog.gemm([act_seg(a1, cout, taps=9)], prog.const_bf16(p + "w2", w2), cout, 9 * cout, B, H, W, cout,
                      bias=prog.const_f32(p + "b2", P(p + "conv2.bias")), resid=x0.t, out_f32=out.t, stats=out.stats)
        return out

    def attnblock(p, x: Act):
        """AttnBlock.forward, unet_ddpm.py:172-197."""
        C, H, W = x.C, x.H, x.W
        T = H * W
        hn = prog.tensor(p + "hn", B * T * C, "bf16")
        prog.gn_apply(src0=x.t, stats0=x.stats, C0=C, P0=x.P, gamma=prog.const_f32(p + "n.w", P(p + "norm.weight")),
                      beta=prog.const_f32(p + "n.b", P(p + "norm.bias")), B=B, H=H, W=W, groups=32, eps=EPS, silu=0,
                      out_bf16=hn)
        o = lower_attention(prog, p + "att", hn, pack_conv1x1(P(p + "q.weight")), pack_conv1x1(P(p + "k.weight")),
                            pack_conv1x1(P(p + "v.weight")), P(p + "q.bias"), P(p + "k.bias"), P(p + "v.bias"), B, T, C,
                            1, int(C) ** (-0.5))
        out = new_act(prog, p + "out", B, C, H, W)
        prog.gemm([act_seg(o, C)], prog.const_bf16(p + "wo", pack_conv1x1(P(p + "proj_out.weight"))), C, C, B, H, W, C,
                  bias=prog.const_f32(p + "bo", P(p + "proj_out.bias")), resid=x.t, out_f32=out.t, stats=out.stats)
        return out

    def resample_conv(p, x: Act, mode):
        """Downsample (pad right/bottom + 3x3 stride 2) / Upsample (nearest x2 + 3x3) on the raw stream."""
        C = x.C
        if mode == "down":
            xb = prog.tensor(p + "xb", B * x.H * x.W * C, "bf16")
            prog.cast(src=x.t, C=C, B=B, H=x.H, W=x.W, out_bf16=xb)
            Ho, Wo = x.H // 2, x.W // 2
            seg = act_seg(xb, C, taps=9, stride=2)
        else:
            Ho, Wo = x.H * 2, x.W * 2
            xb = prog.tensor(p + "xb", B * Ho * Wo * C, "bf16")
            prog.cast(src=x.t, C=C, B=B, H=x.H, W=x.W, out_bf16=xb, resample=1)
            seg = act_seg(xb, C, taps=9)
        out = new_act(prog, p + "out", B, C, Ho, Wo)
        prog.gemm([seg], prog.const_bf16(p + "w", pack_conv3x3(P(p + "conv.weight"))), C, 9 * C, B, Ho, Wo, C,
                  bias=prog.const_f32(p + "b", P(p + "conv.bias")), out_f32=out.t, stats=out.stats)
        return out

    # ---- Model.forward, unet_ddpm.py:305-345 -------------------------------------------------------------
    h0 = new_act(prog, "conv_in.out", B, ch, S, S)
    prog.conv_in_gemm("conv_in", P("conv_in.weight"), P("conv_in.bias"), h0.t, h0.stats, B, S, S, ch)
    hs = [h0]
    res = S
    for lvl in range(nres):
        for b in range(cfg.num_res_blocks):
            h = resblock(f"down.{lvl}.block.{b}.", hs[-1])
            if res in cfg.attn_resolutions:
                h = attnblock(f"down.{lvl}.attn.{b}.", h)
            hs.append(h)
        if lvl != nres - 1:
            hs.append(resample_conv(f"down.{lvl}.downsample.", hs[-1], "down"))
            res //= 2
    h = hs[-1]
    h = resblock("mid.block_1.", h)
    h = attnblock("mid.attn_1.", h)
    h = resblock("mid.block_2.", h)
    for lvl in reversed(range(nres)):
        for b in range(cfg.num_res_blocks + 1):
            h = resblock(f"up.{lvl}.block.{b}.", h, hs.pop())
            if res in cfg.attn_resolutions:
                h = attnblock(f"up.{lvl}.attn.{b}.", h)
        if lvl != 0:
            h = resample_conv(f"up.{lvl}.upsample.", h, "up")
            res *= 2
    assert not hs
    a = prog.tensor("out.a", B * S * S * h.C, "bf16")
    prog.gn_apply(src0=h.t, stats0=h.stats, C0=h.C, P0=h.P, gamma=prog.const_f32("out.n.w", P("norm_out.weight")),
                  beta=prog.const_f32("out.n.b", P("norm_out.bias")), B=B, H=S, W=S, groups=32, eps=EPS, silu=1,
                  out_bf16=a)
    prog.conv_out_gemm("out", a, P("conv_out.weight"), P("conv_out.bias"), B, S, S, h.C, cfg.out_ch)
    prog.meta.update(model="ddpm", out_channels=cfg.out_ch, cond="timestep")
    return prog
