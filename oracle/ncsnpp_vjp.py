"""Oracle for the NEXT scope row: the input-gradient of the DDPM++ score network and of the
Euler-Maruyama purification loop, written out as the kernel-shaped primitives a device backward pass would launch --
test infrastructure only, CPU fp32, nothing in the product imports it.

White-box attacks (eval_sde_adv.py: APGD / EOT through `sdeint_adjoint`, runners/diffpure_sde.py:233-239) need
d(loss)/d(x) only; the weights are frozen, so every layer contributes just its data gradient:

  conv3x3 / conv1x1 / NIN   dgrad = the SAME implicit GEMM with the weights flipped over the taps and transposed
                            (in <-> out): `conv_dgrad`, `nin_dgrad`            -> the existing tcgen05 GEMM kernel
  GroupNorm (+SiLU)         g_y = g * silu'(u), then per (sample, group) the two means  m1 = <g_y gamma>,
                            m2 = <g_y gamma xhat>  and  dx = rstd (g_y gamma - m1 - xhat m2): `gn_silu_vjp`
                                                                                -> a statistics epilogue + one streaming pass
  attention                 dV = P^T dO, dP = dO V^T, dS = P (dP - rowsum(dP P)), dQ = dS K, dK = dS^T Q: `attn_vjp`
                                                                                -> four GEMMs + one row-wise pass
  nearest x2 / 2x2 mean     each other's transpose (x4 resp. /4), concat = split, (x + h)/sqrt2 = g/sqrt2 to both

`vjp()` is checked against torch.autograd on oracle/ncsnpp.forward (itself bit-identical to the reference module,
score_sde/models/ncsnpp.py:232-381). The torchsde adjoint of the reference integrates an augmented SDE backwards with
Brownian replay (package absent offline, unpinned); `purify_sde_vjp` is the exact gradient of the discrete Euler loop
the engine runs (discretise-then-differentiate), which is what a device implementation with stored states would compute.
"""
import math

import torch
import torch.nn.functional as F

from . import ncsnpp as O
from . import sde as OS

INV_SQRT2 = 1.0 / math.sqrt(2.0)


# ---------------------------------------------------------------------------------------------------------
# primitives: forward pieces return what the backward needs; backward pieces use only GEMM / reduction shaped math
# ---------------------------------------------------------------------------------------------------------
def conv_dgrad(g, w):
    """Input gradient of F.conv2d(x, w, padding=k//2): a forward conv with w flipped over the taps and in/out swapped."""
    k = w.shape[-1]
    return F.conv2d(g, w.flip(2, 3).transpose(0, 1), padding=k // 2)


def nin_dgrad(g, W):
    """Input gradient of the NIN contraction y[b,d] = sum_c x[b,c] W[c,d] (layers.py:546-555)."""
    return torch.einsum("bdhw,cd->bchw", g, W)


def silu_grad(u):
    s = torch.sigmoid(u)
    return s * (1 + u * (1 - s))


def gn_stats(x, eps=1e-6):
    """Per (sample, group) mean / rstd with the reference's grouping (min(C/4, 32) groups, layerspp.py:219)."""
    B, C = x.shape[:2]
    G = min(C // 4, 32)
    xg = x.reshape(B, G, -1)
    mean = xg.mean(-1)
    var = xg.var(-1, unbiased=False)
    return mean, torch.rsqrt(var + eps), G


def gn_silu_vjp(x, gamma, g, silu=True, beta=None):
    """dL/dx of y = act(GroupNorm(x)) given g = dL/dy. Two group reductions + one elementwise pass."""
    B, C, H, W = x.shape
    mean, rstd, G = gn_stats(x)
    xhat = ((x.reshape(B, G, -1) - mean[..., None]) * rstd[..., None]).reshape(x.shape)
    gam = gamma[None, :, None, None]
    if silu:
        u = xhat * gam + beta[None, :, None, None]
        g = g * silu_grad(u)
    gx = (g * gam).reshape(B, G, -1)
    xh = xhat.reshape(B, G, -1)
    m1 = gx.mean(-1, keepdim=True)
    m2 = (gx * xh).mean(-1, keepdim=True)
    return (rstd[..., None] * (gx - m1 - xh * m2)).reshape(x.shape)


def up2_vjp(g):      # transpose of nearest x2: sum over each 2x2 block
    n, c, h, w = g.shape
    return g.reshape(n, c, h // 2, 2, w // 2, 2).sum(dim=(3, 5))


def down2_vjp(g):    # transpose of the 2x2 mean: spread g / 4
    return O._up2(g) * 0.25


def attn_vjp(q, k, v, p_mat, g_o, scale):
    """q,k,v,g_o: [B,T,C]; p_mat = softmax(scale q k^T): [B,T,T]. Returns dq, dk, dv."""
    dv = p_mat.transpose(1, 2) @ g_o
    dp = g_o @ v.transpose(1, 2)
    ds = p_mat * (dp - (dp * p_mat).sum(-1, keepdim=True))
    dq = (ds @ k) * scale
    dk = (ds.transpose(1, 2) @ q) * scale
    return dq, dk, dv


# ---------------------------------------------------------------------------------------------------------
# blocks: forward with a tape, backward from the tape
# ---------------------------------------------------------------------------------------------------------
def _res_fwd(sd, p, kw, x, temb):
    t = {"x": x, "kw": kw, "p": p}
    a0 = F.silu(O._gn(x, sd[p + "GroupNorm_0.weight"], sd[p + "GroupNorm_0.bias"]))
    xs = x
    if kw["up"]:
        a0, xs = O._up2(a0), O._up2(x)
    elif kw["down"]:
        a0, xs = O._down2(a0), O._down2(x)
    c0 = F.conv2d(a0, sd[p + "Conv_0.weight"], sd[p + "Conv_0.bias"], padding=1)
    c0 = c0 + F.linear(F.silu(temb), sd[p + "Dense_0.weight"], sd[p + "Dense_0.bias"])[:, :, None, None]
    t["c0"] = c0
    a1 = F.silu(O._gn(c0, sd[p + "GroupNorm_1.weight"], sd[p + "GroupNorm_1.bias"]))
    h = F.conv2d(a1, sd[p + "Conv_1.weight"], sd[p + "Conv_1.bias"], padding=1)
    if (p + "Conv_2.weight") in sd:
        xs = F.conv2d(xs, sd[p + "Conv_2.weight"], sd[p + "Conv_2.bias"])
    return (xs + h) * INV_SQRT2, t


def _res_bwd(sd, t, g):
    p, kw = t["p"], t["kw"]
    g = g * INV_SQRT2                                                   # both branches of (x + h)/sqrt2
    gx = g                                                              # shortcut branch (at the resampled resolution)
    if (p + "Conv_2.weight") in sd:
        gx = conv_dgrad(gx, sd[p + "Conv_2.weight"])
    ga1 = conv_dgrad(g, sd[p + "Conv_1.weight"])
    gc0 = gn_silu_vjp(t["c0"], sd[p + "GroupNorm_1.weight"], ga1, True, sd[p + "GroupNorm_1.bias"])
    ga0 = conv_dgrad(gc0, sd[p + "Conv_0.weight"])                      # (the temb add has no x dependence)
    if kw["up"]:
        ga0, gx = up2_vjp(ga0), up2_vjp(gx)
    elif kw["down"]:
        ga0, gx = down2_vjp(ga0), down2_vjp(gx)
    return gx + gn_silu_vjp(t["x"], sd[p + "GroupNorm_0.weight"], ga0, True, sd[p + "GroupNorm_0.bias"])


def _attn_fwd(sd, p, x):
    B, C, H, W = x.shape
    hn = O._gn(x, sd[p + "GroupNorm_0.weight"], sd[p + "GroupNorm_0.bias"])
    q = O._nin(hn, sd[p + "NIN_0.W"], sd[p + "NIN_0.b"])
    k = O._nin(hn, sd[p + "NIN_1.W"], sd[p + "NIN_1.b"])
    v = O._nin(hn, sd[p + "NIN_2.W"], sd[p + "NIN_2.b"])
    tok = lambda z: z.reshape(B, C, H * W).transpose(1, 2)              # noqa: E731  [B,T,C]
    scale = int(C) ** (-0.5)
    pm = torch.softmax(scale * tok(q) @ tok(k).transpose(1, 2), dim=-1)
    o = (pm @ tok(v)).transpose(1, 2).reshape(B, C, H, W)
    h = O._nin(o, sd[p + "NIN_3.W"], sd[p + "NIN_3.b"])
    return (x + h) * INV_SQRT2, {"x": x, "p": p, "q": tok(q), "k": tok(k), "v": tok(v), "pm": pm, "scale": scale}


def _attn_bwd(sd, t, g):
    p, x = t["p"], t["x"]
    B, C, H, W = x.shape
    g = g * INV_SQRT2
    go = nin_dgrad(g, sd[p + "NIN_3.W"]).reshape(B, C, H * W).transpose(1, 2)
    dq, dk, dv = attn_vjp(t["q"], t["k"], t["v"], t["pm"], go, t["scale"])
    img = lambda z: z.transpose(1, 2).reshape(B, C, H, W)               # noqa: E731
    ghn = nin_dgrad(img(dq), sd[p + "NIN_0.W"]) + nin_dgrad(img(dk), sd[p + "NIN_1.W"]) + \
        nin_dgrad(img(dv), sd[p + "NIN_2.W"])
    return g + gn_silu_vjp(x, sd[p + "GroupNorm_0.weight"], ghn, False)


def forward_with_tape(cfg, sd, x, time_cond):
    """oracle/ncsnpp.forward, recording per block what its backward needs."""
    mods = O.module_list(cfg)
    nres = len(cfg.ch_mult)
    idx = 0
    tape = []

    def nxt():
        nonlocal idx
        kind, kw = mods[idx]
        p = f"all_modules.{idx}."
        idx += 1
        return kind, kw, p

    def res(h):
        _, kw, p = nxt()
        h, t = _res_fwd(sd, p, kw, h, temb)
        tape.append(("res", t))
        return h

    def attn(h):
        _, kw, p = nxt()
        h, t = _attn_fwd(sd, p, h)
        tape.append(("attn", t))
        return h

    temb = O.timestep_embedding(time_cond, cfg.nf)
    _, _, p = nxt()
    temb = F.linear(temb, sd[p + "weight"], sd[p + "bias"])
    _, _, p = nxt()
    temb = F.linear(F.silu(temb), sd[p + "weight"], sd[p + "bias"])
    _, _, p = nxt()
    tape.append(("conv_in", {"p": p}))
    hs = [F.conv2d(x, sd[p + "weight"], sd[p + "bias"], padding=1)]
    tape.append(("push", None))
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            h = res(hs[-1])
            if h.shape[-1] in cfg.attn_resolutions:
                h = attn(h)
            hs.append(h)
            tape.append(("push", None))
        if lvl != nres - 1:
            hs.append(res(hs[-1]))
            tape.append(("push", None))
    h = hs[-1]
    h = res(h)
    h = attn(h)
    h = res(h)
    for lvl in reversed(range(nres)):
        for _ in range(cfg.num_res_blocks + 1):
            skip = hs.pop()
            tape.append(("cat", {"c": h.shape[1]}))
            h = res(torch.cat([h, skip], dim=1))
        if h.shape[-1] in cfg.attn_resolutions:
            h = attn(h)
        if lvl != 0:
            h = res(h)
    _, _, p = nxt()
    tape.append(("gn_out", {"p": p, "x": h}))
    h = F.silu(O._gn(h, sd[p + "weight"], sd[p + "bias"]))
    _, _, p = nxt()
    tape.append(("conv_out", {"p": p}))
    return F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1), tape


def vjp(cfg, sd, x, time_cond, g_out):
    """dL/dx for L = <g_out, forward(cfg, sd, x, time_cond)>."""
    _, tape = forward_with_tape(cfg, sd, x, time_cond)
    g = g_out
    # Walk the tape backwards; `g` is the gradient of the main stream. The up path consumed the skip stack last-in
    # first-out, so walking it backwards meets the skip tensors in PUSH order: every 'cat' splits the skip part off and
    # appends it to `skips`; every 'push' marker (met in reverse push order on the way down) takes the latest one back.
    skips = []
    for kind, t in reversed(tape):
        if kind == "conv_out":
            g = conv_dgrad(g, sd[t["p"] + "weight"])
        elif kind == "gn_out":
            g = gn_silu_vjp(t["x"], sd[t["p"] + "weight"], g, True, sd[t["p"] + "bias"])
        elif kind == "res":
            g = _res_bwd(sd, t, g)
        elif kind == "attn":
            g = _attn_bwd(sd, t, g)
        elif kind == "cat":
            skips.append(g[:, t["c"]:])
            g = g[:, :t["c"]]
        elif kind == "push":
            g = g + skips.pop()
        elif kind == "conv_in":
            g = conv_dgrad(g, sd[t["p"] + "weight"])
    assert not skips
    return g


# ---------------------------------------------------------------------------------------------------------
# loop level: exact gradient of the discrete Euler-Maruyama loop (oracle/sde.py:purify_sde)
# ---------------------------------------------------------------------------------------------------------
def purify_sde_vjp(cfg, sd, x0, t_star, init_noise, step_noise, g_out):
    """dL/dx0 for L = <g_out, purify_sde(unet, x0, ...)> with score_type 'score_sde': forward stores the states x_k,
    backward runs lambda_k = (1 + beta_k h_k / 2) lambda_{k+1} + J_k^T [-(beta_k/sigma_k) h_k lambda_{k+1}] ."""
    grid = OS.time_grid(t_star)                                          # the loop of oracle/sde.py:purify_sde, as
    t, h = grid[:-1], grid[1:] - grid[:-1]                               # x <- c0 x + c1 eps + c2 z per step
    s = 1 - t
    beta = OS.BETA_MIN + s * (OS.BETA_MAX - OS.BETA_MIN)
    sigma = torch.sqrt(1. - torch.exp(2. * (-0.25 * s ** 2 * (OS.BETA_MAX - OS.BETA_MIN) - 0.5 * s * OS.BETA_MIN)))
    cond = s * 999
    coef = torch.stack([1 + 0.5 * beta * h, -(beta / sigma) * h, torch.sqrt(beta) * torch.sqrt(h)], 1)
    betas = torch.linspace(OS.BETA_MIN / OS.N_SCALES, OS.BETA_MAX / OS.N_SCALES, OS.N_SCALES).float()
    a = (1 - betas).cumprod(dim=0)
    sx, se = float(a[t_star - 1].sqrt()), float((1.0 - a[t_star - 1]).sqrt())
    B = x0.shape[0]
    xs = [sx * x0 + se * init_noise]
    with torch.no_grad():
        for k in range(len(cond)):
            eps = O.forward(cfg, sd, xs[-1], torch.full((B,), float(cond[k])))
            xs.append(float(coef[k, 0]) * xs[-1] + float(coef[k, 1]) * eps + float(coef[k, 2]) * step_noise[k])
        lam = g_out
        for k in reversed(range(len(cond))):
            lam = float(coef[k, 0]) * lam + vjp(cfg, sd, xs[k], torch.full((B,), float(cond[k])), float(coef[k, 1]) * lam)
    return sx * lam, xs[-1]


__all__ = ["conv_dgrad", "nin_dgrad", "gn_silu_vjp", "attn_vjp", "up2_vjp", "down2_vjp", "forward_with_tape", "vjp",
           "purify_sde_vjp", "OS"]
