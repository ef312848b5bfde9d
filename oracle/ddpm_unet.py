"""Oracle restatement of the CelebA-HQ DDPM UNet (CPU, fp32) -- test infrastructure only.

Follows ddpm/unet_ddpm.py:200-345 (`Model`), ResnetBlock L85-142, AttnBlock L145-197, Up/Downsample L44-82,
get_timestep_embedding L14-32, GroupNorm(32, eps=1e-6) L40-41. Parameter names are the reference's.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

CELEBA_CFG = SimpleNamespace(image_size=256, ch=128, out_ch=3, ch_mult=(1, 1, 2, 2, 4, 4), num_res_blocks=2,
                             attn_resolutions=(16,), in_channels=3)  # configs/celeba.yml:13-25


def tiny_cfg(image_size=32, ch=64, ch_mult=(1, 2, 2), num_res_blocks=1, attn_resolutions=(16,)):
    return SimpleNamespace(image_size=image_size, ch=ch, out_ch=3, ch_mult=tuple(ch_mult),
                           num_res_blocks=num_res_blocks, attn_resolutions=tuple(attn_resolutions), in_channels=3)


def _res_shapes(sh, p, cin, cout, temb):
    sh[p + "norm1.weight"] = (cin,); sh[p + "norm1.bias"] = (cin,)
    sh[p + "conv1.weight"] = (cout, cin, 3, 3); sh[p + "conv1.bias"] = (cout,)
    sh[p + "temb_proj.weight"] = (cout, temb); sh[p + "temb_proj.bias"] = (cout,)
    sh[p + "norm2.weight"] = (cout,); sh[p + "norm2.bias"] = (cout,)
    sh[p + "conv2.weight"] = (cout, cout, 3, 3); sh[p + "conv2.bias"] = (cout,)
    if cin != cout:
        sh[p + "nin_shortcut.weight"] = (cout, cin, 1, 1); sh[p + "nin_shortcut.bias"] = (cout,)


def _attn_shapes(sh, p, c):
    sh[p + "norm.weight"] = (c,); sh[p + "norm.bias"] = (c,)
    for n in ("q", "k", "v", "proj_out"):
        sh[p + n + ".weight"] = (c, c, 1, 1); sh[p + n + ".bias"] = (c,)


def param_shapes(cfg):
    """Registration order of ddpm/unet_ddpm.py:216-303."""
    sh = {}
    ch, temb = cfg.ch, cfg.ch * 4
    sh["temb.dense.0.weight"] = (temb, ch); sh["temb.dense.0.bias"] = (temb,)
    sh["temb.dense.1.weight"] = (temb, temb); sh["temb.dense.1.bias"] = (temb,)
    sh["conv_in.weight"] = (ch, cfg.in_channels, 3, 3); sh["conv_in.bias"] = (ch,)
    nres = len(cfg.ch_mult)
    in_mult = (1,) + tuple(cfg.ch_mult)
    res = cfg.image_size
    block_in = None
    for lvl in range(nres):
        block_in, block_out = ch * in_mult[lvl], ch * cfg.ch_mult[lvl]
        for b in range(cfg.num_res_blocks):
            _res_shapes(sh, f"down.{lvl}.block.{b}.", block_in, block_out, temb)
            block_in = block_out
        if res in cfg.attn_resolutions:
            for b in range(cfg.num_res_blocks):
                _attn_shapes(sh, f"down.{lvl}.attn.{b}.", block_in)
        if lvl != nres - 1:
            sh[f"down.{lvl}.downsample.conv.weight"] = (block_in, block_in, 3, 3)
            sh[f"down.{lvl}.downsample.conv.bias"] = (block_in,)
            res //= 2
    _res_shapes(sh, "mid.block_1.", block_in, block_in, temb)
    _attn_shapes(sh, "mid.attn_1.", block_in)
    _res_shapes(sh, "mid.block_2.", block_in, block_in, temb)
    ups = {}
    for lvl in reversed(range(nres)):
        block_out = ch * cfg.ch_mult[lvl]
        skip_in = ch * cfg.ch_mult[lvl]
        d = {}
        for b in range(cfg.num_res_blocks + 1):
            if b == cfg.num_res_blocks:
                skip_in = ch * in_mult[lvl]
            _res_shapes(d, f"up.{lvl}.block.{b}.", block_in + skip_in, block_out, temb)
            block_in = block_out
        if res in cfg.attn_resolutions:
            for b in range(cfg.num_res_blocks + 1):
                _attn_shapes(d, f"up.{lvl}.attn.{b}.", block_in)
        if lvl != 0:
            d[f"up.{lvl}.upsample.conv.weight"] = (block_in, block_in, 3, 3)
            d[f"up.{lvl}.upsample.conv.bias"] = (block_in,)
            res *= 2
        ups[lvl] = d
    for lvl in range(nres):          # `self.up.insert(0, up)` -> registered in ascending level order
        sh.update(ups[lvl])
    sh["norm_out.weight"] = (block_in,); sh["norm_out.bias"] = (block_in,)
    sh["conv_out.weight"] = (cfg.out_ch, block_in, 3, 3); sh["conv_out.bias"] = (cfg.out_ch,)
    return sh


def timestep_embedding(t, dim):                                   # unet_ddpm.py:14-32
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -emb)
    emb = t.float()[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)


def _swish(x):
    return x * torch.sigmoid(x)


def _gn(x, sd, p):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def _res(sd, p, x, temb):                                          # unet_ddpm.py:123-142
    h = F.conv2d(_swish(_gn(x, sd, p + "norm1")), sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    h = h + F.linear(_swish(temb), sd[p + "temb_proj.weight"], sd[p + "temb_proj.bias"])[:, :, None, None]
    h = F.conv2d(_swish(_gn(h, sd, p + "norm2")), sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if (p + "nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + "nin_shortcut.weight"], sd[p + "nin_shortcut.bias"])
    return x + h


def _attn(sd, p, x):                                               # unet_ddpm.py:172-197
    h_ = _gn(x, sd, p + "norm")
    q = F.conv2d(h_, sd[p + "q.weight"], sd[p + "q.bias"])
    k = F.conv2d(h_, sd[p + "k.weight"], sd[p + "k.bias"])
    v = F.conv2d(h_, sd[p + "v.weight"], sd[p + "v.bias"])
    b, c, h, w = q.shape
    w_ = torch.bmm(q.reshape(b, c, h * w).permute(0, 2, 1), k.reshape(b, c, h * w)) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    h_ = torch.bmm(v.reshape(b, c, h * w), w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + F.conv2d(h_, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])


def forward(cfg, sd, x, t):
    """Model.forward, unet_ddpm.py:305-345. t: [B] integer timesteps."""
    nres = len(cfg.ch_mult)
    temb = timestep_embedding(t, cfg.ch)
    temb = F.linear(temb, sd["temb.dense.0.weight"], sd["temb.dense.0.bias"])
    temb = F.linear(_swish(temb), sd["temb.dense.1.weight"], sd["temb.dense.1.bias"])
    hs = [F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)]
    for lvl in range(nres):
        for b in range(cfg.num_res_blocks):
            h = _res(sd, f"down.{lvl}.block.{b}.", hs[-1], temb)
            if f"down.{lvl}.attn.{b}.norm.weight" in sd:
                h = _attn(sd, f"down.{lvl}.attn.{b}.", h)
            hs.append(h)
        if lvl != nres - 1:
            xp = F.pad(hs[-1], (0, 1, 0, 1))                        # unet_ddpm.py:76-79
            hs.append(F.conv2d(xp, sd[f"down.{lvl}.downsample.conv.weight"], sd[f"down.{lvl}.downsample.conv.bias"],
                               stride=2))
    h = hs[-1]
    h = _res(sd, "mid.block_1.", h, temb)
    h = _attn(sd, "mid.attn_1.", h)
    h = _res(sd, "mid.block_2.", h, temb)
    for lvl in reversed(range(nres)):
        for b in range(cfg.num_res_blocks + 1):
            h = _res(sd, f"up.{lvl}.block.{b}.", torch.cat([h, hs.pop()], dim=1), temb)
            if f"up.{lvl}.attn.{b}.norm.weight" in sd:
                h = _attn(sd, f"up.{lvl}.attn.{b}.", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"up.{lvl}.upsample.conv.weight"], sd[f"up.{lvl}.upsample.conv.bias"], padding=1)
    h = _swish(_gn(h, sd, "norm_out"))
    return F.conv2d(h, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
