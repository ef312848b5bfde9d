"""Oracle restatement of the guided_diffusion (ADM) UNet (CPU, fp32) -- test infrastructure only.

Follows guided_diffusion/unet.py:404-671 (`UNetModel`) for DiffPure's ImageNet configuration
(configs/imagenet.yml:5-19 over script_util.py:51-73,136-192): ResBlock with scale-shift norm and
resblock_updown (L151-264), AttentionBlock + QKVAttentionLegacy (L267-362), GroupNorm32 (nn.py:25-27, eps 1e-5),
timestep_embedding (nn.py:111-129, [cos | sin], /half), learn_sigma (6 output channels).
`use_fp16` is a storage/compute precision choice of the reference (unet.py:626-632); this restatement computes in
fp32 -- `forward(..., fp16_torso=True)` emulates the reference's casts.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

IMAGENET_CFG = SimpleNamespace(image_size=256, model_channels=256, out_channels=6, num_res_blocks=2,
                               channel_mult=(1, 1, 2, 2, 4, 4), attention_ds=(8, 16, 32), num_head_channels=64)


def tiny_cfg(image_size=64, model_channels=64, channel_mult=(1, 2, 3, 4), num_res_blocks=1,
             attention_resolutions=(32, 16, 8)):
    return SimpleNamespace(image_size=image_size, model_channels=model_channels, out_channels=6,
                           num_res_blocks=num_res_blocks, channel_mult=tuple(channel_mult),
                           attention_ds=tuple(image_size // r for r in attention_resolutions), num_head_channels=64)


def block_plan(cfg):
    """input_blocks / middle / output_blocks as lists of layer records (unet.py:486-606)."""
    mc = cfg.model_channels
    ch = int(cfg.channel_mult[0] * mc)
    inp = [[("conv_in", dict(cout=ch))]]
    chans = [ch]
    ds = 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", dict(cin=ch, cout=int(mult * mc), up=False, down=False))]
            ch = int(mult * mc)
            if ds in cfg.attention_ds:
                layers.append(("attn", dict(c=ch)))
            inp.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inp.append([("res", dict(cin=ch, cout=ch, up=False, down=True))])
            chans.append(ch)
            ds *= 2
    mid = [("res", dict(cin=ch, cout=ch, up=False, down=False)), ("attn", dict(c=ch)),
           ("res", dict(cin=ch, cout=ch, up=False, down=False))]
    out = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", dict(cin=ch + ich, cout=int(mc * mult), up=False, down=False))]
            ch = int(mc * mult)
            if ds in cfg.attention_ds:
                layers.append(("attn", dict(c=ch)))
            if level and i == cfg.num_res_blocks:
                layers.append(("res", dict(cin=ch, cout=ch, up=True, down=False)))
                ds //= 2
            out.append(layers)
    return inp, mid, out, ch


def _layer_shapes(sh, p, kind, kw, emb):
    if kind == "conv_in":
        sh[p + "weight"] = (kw["cout"], 3, 3, 3); sh[p + "bias"] = (kw["cout"],)
    elif kind == "res":
        cin, cout = kw["cin"], kw["cout"]
        sh[p + "in_layers.0.weight"] = (cin,); sh[p + "in_layers.0.bias"] = (cin,)
        sh[p + "in_layers.2.weight"] = (cout, cin, 3, 3); sh[p + "in_layers.2.bias"] = (cout,)
        sh[p + "emb_layers.1.weight"] = (2 * cout, emb); sh[p + "emb_layers.1.bias"] = (2 * cout,)
        sh[p + "out_layers.0.weight"] = (cout,); sh[p + "out_layers.0.bias"] = (cout,)
        sh[p + "out_layers.3.weight"] = (cout, cout, 3, 3); sh[p + "out_layers.3.bias"] = (cout,)
        if cin != cout:
            sh[p + "skip_connection.weight"] = (cout, cin, 1, 1); sh[p + "skip_connection.bias"] = (cout,)
    elif kind == "attn":
        c = kw["c"]
        sh[p + "norm.weight"] = (c,); sh[p + "norm.bias"] = (c,)
        sh[p + "qkv.weight"] = (3 * c, c, 1); sh[p + "qkv.bias"] = (3 * c,)
        sh[p + "proj_out.weight"] = (c, c, 1); sh[p + "proj_out.bias"] = (c,)


def param_shapes(cfg):
    inp, mid, out, ch = block_plan(cfg)
    emb = cfg.model_channels * 4
    sh = {"time_embed.0.weight": (emb, cfg.model_channels), "time_embed.0.bias": (emb,),
          "time_embed.2.weight": (emb, emb), "time_embed.2.bias": (emb,)}
    for i, layers in enumerate(inp):
        for j, (kind, kw) in enumerate(layers):
            _layer_shapes(sh, f"input_blocks.{i}.{j}.", kind, kw, emb)
    for j, (kind, kw) in enumerate(mid):
        _layer_shapes(sh, f"middle_block.{j}.", kind, kw, emb)
    for i, layers in enumerate(out):
        for j, (kind, kw) in enumerate(layers):
            _layer_shapes(sh, f"output_blocks.{i}.{j}.", kind, kw, emb)
    sh["out.0.weight"] = (ch,); sh["out.0.bias"] = (ch,)
    sh["out.2.weight"] = (cfg.out_channels, ch, 3, 3); sh["out.2.bias"] = (cfg.out_channels,)
    return sh


def timestep_embedding(t, dim, max_period=10000):                  # nn.py:111-129
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x, sd, p):                                                  # nn.py:25-27 (fp32 compute, eps 1e-5)
    return F.group_norm(x.float(), 32, sd[p + ".weight"].float(), sd[p + ".bias"].float(), 1e-5).type(x.dtype)


def _res(sd, p, kw, x, emb, cd):                                    # unet.py:244-264
    w = lambda n: sd[p + n].to(cd)  # noqa: E731
    h = F.silu(_gn(x, sd, p + "in_layers.0"))
    if kw["up"]:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif kw["down"]:
        h = F.avg_pool2d(h, 2, 2)
        x = F.avg_pool2d(x, 2, 2)
    h = F.conv2d(h, w("in_layers.2.weight"), w("in_layers.2.bias"), padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"]).type(h.dtype)
    scale, shift = torch.chunk(emb_out[:, :, None, None], 2, dim=1)
    h = _gn(h, sd, p + "out_layers.0") * (1 + scale) + shift
    h = F.conv2d(F.silu(h), w("out_layers.3.weight"), w("out_layers.3.bias"), padding=1)
    if (p + "skip_connection.weight") in sd:
        x = F.conv2d(x, w("skip_connection.weight"), w("skip_connection.bias"))
    return x + h


def _attn(sd, p, x, heads_ch, cd):                                  # unet.py:307-313,345-362
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(xf, sd, p + "norm"), sd[p + "qkv.weight"].to(cd), sd[p + "qkv.bias"].to(cd))
    n_heads = c // heads_ch
    bs, width, length = qkv.shape
    ch = width // (3 * n_heads)
    q, k, v = qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    weight = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    weight = torch.softmax(weight.float(), dim=-1).type(weight.dtype)
    a = torch.einsum("bts,bcs->bct", weight, v).reshape(bs, -1, length)
    h = F.conv1d(a, sd[p + "proj_out.weight"].to(cd), sd[p + "proj_out.bias"].to(cd))
    return (xf + h).reshape(b, c, hh, ww)


def forward(cfg, sd, x, timesteps, fp16_torso=False):
    """UNetModel.forward, unet.py:642-671. Returns [B, 6, H, W] fp32."""
    inp, mid, out, _ = block_plan(cfg)
    cd = torch.float16 if fp16_torso else torch.float32
    emb = timestep_embedding(timesteps, cfg.model_channels)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])

    def run(prefix, layers, h):
        for j, (kind, kw) in enumerate(layers):
            p = f"{prefix}{j}."
            if kind == "conv_in":
                h = F.conv2d(h, sd[p + "weight"].to(cd), sd[p + "bias"].to(cd), padding=1)
            elif kind == "res":
                h = _res(sd, p, kw, h, emb, cd)
            else:
                h = _attn(sd, p, h, cfg.num_head_channels, cd)
        return h

    hs = []
    h = x.type(cd)
    for i, layers in enumerate(inp):
        h = run(f"input_blocks.{i}.", layers, h)
        hs.append(h)
    h = run("middle_block.", mid, h)
    for i, layers in enumerate(out):
        h = run(f"output_blocks.{i}.", layers, torch.cat([h, hs.pop()], dim=1))
    h = h.type(x.dtype)
    h = F.silu(_gn(h, sd, "out.0"))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)
