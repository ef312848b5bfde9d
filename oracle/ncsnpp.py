"""Oracle restatement of the DDPM++ / NCSN++ score network (CPU, fp32) -- test infrastructure only.

Follows score_sde/models/ncsnpp.py:35-381 for the configuration DiffPure uses (configs/cifar10.yml:18-40:
biggan res-blocks, skip_rescale, positional embedding, conditional, progressive none, fir False, swish).
Parameter names are the reference's (`all_modules.<i>.<name>`), so a reference state_dict loads as is.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

CIFAR10_CFG = SimpleNamespace(  # configs/cifar10.yml:1-40
    image_size=32, num_channels=3, nf=128, ch_mult=(1, 2, 2, 2), num_res_blocks=8,
    attn_resolutions=(16,), num_scales=1000, beta_min=0.1, beta_max=20.0)


def tiny_cfg(nf=64, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(8,), image_size=16):
    """A reduced configuration with the same block types, for fast tests."""
    return SimpleNamespace(image_size=image_size, num_channels=3, nf=nf, ch_mult=tuple(ch_mult),
                           num_res_blocks=num_res_blocks, attn_resolutions=tuple(attn_resolutions),
                           num_scales=1000, beta_min=0.1, beta_max=20.0)


def module_list(cfg):
    """The `all_modules` sequence of ncsnpp.py:68-230 as (kind, kwargs) records, in construction order."""
    nf, ch_mult, nrb = cfg.nf, cfg.ch_mult, cfg.num_res_blocks
    nres = len(ch_mult)
    all_res = [cfg.image_size // (2 ** i) for i in range(nres)]
    mods = [("linear", dict(cin=nf, cout=4 * nf)), ("linear", dict(cin=4 * nf, cout=4 * nf)),
            ("conv3x3", dict(cin=cfg.num_channels, cout=nf))]
    hs_c = [nf]
    in_ch = nf
    for lvl in range(nres):                                              # ncsnpp.py:155-183
        for _ in range(nrb):
            out_ch = nf * ch_mult[lvl]
            mods.append(("res", dict(cin=in_ch, cout=out_ch, up=False, down=False)))
            in_ch = out_ch
            if all_res[lvl] in cfg.attn_resolutions:
                mods.append(("attn", dict(c=in_ch)))
            hs_c.append(in_ch)
        if lvl != nres - 1:
            mods.append(("res", dict(cin=in_ch, cout=in_ch, up=False, down=True)))
            hs_c.append(in_ch)
    in_ch = hs_c[-1]                                                     # ncsnpp.py:185-188
    mods += [("res", dict(cin=in_ch, cout=in_ch, up=False, down=False)), ("attn", dict(c=in_ch)),
             ("res", dict(cin=in_ch, cout=in_ch, up=False, down=False))]
    for lvl in reversed(range(nres)):                                    # ncsnpp.py:191-222
        for _ in range(nrb + 1):
            out_ch = nf * ch_mult[lvl]
            mods.append(("res", dict(cin=in_ch + hs_c.pop(), cout=out_ch, up=False, down=False)))
            in_ch = out_ch
        if all_res[lvl] in cfg.attn_resolutions:
            mods.append(("attn", dict(c=in_ch)))
        if lvl != 0:
            mods.append(("res", dict(cin=in_ch, cout=in_ch, up=True, down=False)))
    assert not hs_c
    mods += [("gn", dict(c=in_ch)), ("conv3x3", dict(cin=in_ch, cout=cfg.num_channels))]  # ncsnpp.py:225-228
    return mods


def param_shapes(cfg):
    """Ordered name -> shape of every parameter, matching the reference state_dict."""
    shapes = {}
    temb = 4 * cfg.nf
    for i, (kind, kw) in enumerate(module_list(cfg)):
        p = f"all_modules.{i}."
        if kind == "linear":
            shapes[p + "weight"] = (kw["cout"], kw["cin"])
            shapes[p + "bias"] = (kw["cout"],)
        elif kind == "conv3x3":
            shapes[p + "weight"] = (kw["cout"], kw["cin"], 3, 3)
            shapes[p + "bias"] = (kw["cout"],)
        elif kind == "gn":
            shapes[p + "weight"] = (kw["c"],)
            shapes[p + "bias"] = (kw["c"],)
        elif kind == "res":                                              # layerspp.py:212-240
            cin, cout = kw["cin"], kw["cout"]
            shapes[p + "GroupNorm_0.weight"] = (cin,)
            shapes[p + "GroupNorm_0.bias"] = (cin,)
            shapes[p + "Conv_0.weight"] = (cout, cin, 3, 3)
            shapes[p + "Conv_0.bias"] = (cout,)
            shapes[p + "Dense_0.weight"] = (cout, temb)
            shapes[p + "Dense_0.bias"] = (cout,)
            shapes[p + "GroupNorm_1.weight"] = (cout,)
            shapes[p + "GroupNorm_1.bias"] = (cout,)
            shapes[p + "Conv_1.weight"] = (cout, cout, 3, 3)
            shapes[p + "Conv_1.bias"] = (cout,)
            if cin != cout or kw["up"] or kw["down"]:
                shapes[p + "Conv_2.weight"] = (cout, cin, 1, 1)
                shapes[p + "Conv_2.bias"] = (cout,)
        elif kind == "attn":                                             # layerspp.py:65-73
            c = kw["c"]
            shapes[p + "GroupNorm_0.weight"] = (c,)
            shapes[p + "GroupNorm_0.bias"] = (c,)
            for j in range(4):
                shapes[p + f"NIN_{j}.W"] = (c, c)
                shapes[p + f"NIN_{j}.b"] = (c,)
    return shapes


def timestep_embedding(timesteps, dim, max_positions=10000):
    """layers.py:515-529 -- [sin | cos], freq_i = exp(-ln(max_pos) * i / (half - 1))."""
    half = dim // 2
    emb = math.log(max_positions) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32, device=timesteps.device) * -emb)
    emb = timesteps.float()[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)


def _gn(x, w, b, eps=1e-6):
    c = x.shape[1]
    return F.group_norm(x, min(c // 4, 32), w, b, eps)                   # layerspp.py:219 (eps 1e-6)


def _nin(x, W, b):
    """layers.py:546-555: contraction over channels with W [in, out]."""
    y = torch.einsum("bchw,cd->bdhw", x, W)
    return y + b[None, :, None, None]


def _up2(x):                                                             # up_or_down_sampling.py:67-71
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def _down2(x):                                                           # up_or_down_sampling.py:74-77
    n, c, h, w = x.shape
    return x.reshape(n, c, h // 2, 2, w // 2, 2).mean(dim=(3, 5))


def _res(sd, p, kw, x, temb):
    """ResnetBlockBigGANpp.forward, layerspp.py:242-274."""
    h = F.silu(_gn(x, sd[p + "GroupNorm_0.weight"], sd[p + "GroupNorm_0.bias"]))
    if kw["up"]:
        h, x = _up2(h), _up2(x)
    elif kw["down"]:
        h, x = _down2(h), _down2(x)
    h = F.conv2d(h, sd[p + "Conv_0.weight"], sd[p + "Conv_0.bias"], padding=1)
    h = h + F.linear(F.silu(temb), sd[p + "Dense_0.weight"], sd[p + "Dense_0.bias"])[:, :, None, None]
    h = F.silu(_gn(h, sd[p + "GroupNorm_1.weight"], sd[p + "GroupNorm_1.bias"]))
    h = F.conv2d(h, sd[p + "Conv_1.weight"], sd[p + "Conv_1.bias"], padding=1)
    if (p + "Conv_2.weight") in sd:
        x = F.conv2d(x, sd[p + "Conv_2.weight"], sd[p + "Conv_2.bias"])
    return (x + h) / math.sqrt(2.0)


def _attn(sd, p, x):
    """AttnBlockpp.forward, layerspp.py:75-91 (skip_rescale True)."""
    B, C, H, W = x.shape
    h = _gn(x, sd[p + "GroupNorm_0.weight"], sd[p + "GroupNorm_0.bias"])
    q = _nin(h, sd[p + "NIN_0.W"], sd[p + "NIN_0.b"])
    k = _nin(h, sd[p + "NIN_1.W"], sd[p + "NIN_1.b"])
    v = _nin(h, sd[p + "NIN_2.W"], sd[p + "NIN_2.b"])
    w = torch.einsum("bchw,bcij->bhwij", q, k) * (int(C) ** (-0.5))
    w = F.softmax(w.reshape(B, H, W, H * W), dim=-1).reshape(B, H, W, H, W)
    h = torch.einsum("bhwij,bcij->bchw", w, v)
    h = _nin(h, sd[p + "NIN_3.W"], sd[p + "NIN_3.b"])
    return (x + h) / math.sqrt(2.0)


def forward(cfg, sd, x, time_cond):
    """NCSNpp.forward, ncsnpp.py:232-381. x: [B,3,H,W] in [-1,1]; time_cond: [B] float labels (= 999*t)."""
    mods = module_list(cfg)
    nres = len(cfg.ch_mult)
    idx = 0

    def nxt():
        nonlocal idx
        kind, kw = mods[idx]
        p = f"all_modules.{idx}."
        idx += 1
        return kind, kw, p

    temb = timestep_embedding(time_cond, cfg.nf)                         # ncsnpp.py:246
    _, _, p = nxt()
    temb = F.linear(temb, sd[p + "weight"], sd[p + "bias"])
    _, _, p = nxt()
    temb = F.linear(F.silu(temb), sd[p + "weight"], sd[p + "bias"])
    _, _, p = nxt()
    hs = [F.conv2d(x, sd[p + "weight"], sd[p + "bias"], padding=1)]      # ncsnpp.py:268
    for lvl in range(nres):
        for _ in range(cfg.num_res_blocks):
            _, kw, p = nxt()
            h = _res(sd, p, kw, hs[-1], temb)
            if h.shape[-1] in cfg.attn_resolutions:
                _, kw, p = nxt()
                h = _attn(sd, p, h)
            hs.append(h)
        if lvl != nres - 1:
            _, kw, p = nxt()
            hs.append(_res(sd, p, kw, hs[-1], temb))
    h = hs[-1]
    _, kw, p = nxt()
    h = _res(sd, p, kw, h, temb)
    _, kw, p = nxt()
    h = _attn(sd, p, h)
    _, kw, p = nxt()
    h = _res(sd, p, kw, h, temb)
    for lvl in reversed(range(nres)):
        for _ in range(cfg.num_res_blocks + 1):
            _, kw, p = nxt()
            h = _res(sd, p, kw, torch.cat([h, hs.pop()], dim=1), temb)
        if h.shape[-1] in cfg.attn_resolutions:
            _, kw, p = nxt()
            h = _attn(sd, p, h)
        if lvl != 0:
            _, kw, p = nxt()
            h = _res(sd, p, kw, h, temb)
    assert not hs
    _, _, p = nxt()
    h = F.silu(_gn(h, sd[p + "weight"], sd[p + "bias"]))
    _, _, p = nxt()
    h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
    assert idx == len(mods)
    return h
