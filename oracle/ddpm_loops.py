"""Oracle restatement of the two DDPM ancestral purification loops (CPU, fp32) -- test infrastructure only.

guided:  runners/diffpure_guided.py:41-89 driving GaussianDiffusion.p_sample (gaussian_diffusion.py:403-447) ->
         p_mean_variance (L240-334, LEARNED_RANGE variance L277-284, eps prediction with x0 clamp L305,317-322)
         through SpacedDiffusion/_WrappedModel (respace.py:71-136; timestep_respacing '1000', rescale_timesteps).
celeba:  runners/diffpure_ddpm.py:99-142 with image_editing_denoising_step_flexible_mask (L37-54).
Noise is injected: `step_noise[k]` is the k-th torch.randn_like(x) the reference draws (one per step including the
masked t = 0 step), `init_noise` the forward-diffusion draw.
"""
import numpy as np
import torch


def _extract(arr, t, shape):
    """gaussian_diffusion.py:903-916: float64 table -> device, gather, cast to fp32, broadcast."""
    res = torch.from_numpy(arr)[t].float()
    while len(res.shape) < len(shape):
        res = res[..., None]
    return res.expand(shape)


def _kept_steps(n, spec):
    """respace.py:7-60 (space_timesteps): which base timesteps a `timestep_respacing` spec keeps, ascending."""
    if not spec:
        return list(range(n))
    if isinstance(spec, str) and spec.startswith("ddim"):
        k = int(spec[len("ddim"):])
        hits = [st for st in range(1, n) if len(range(0, n, st)) == k]
        if not hits:
            raise ValueError("no integer stride gives that many steps")
        return list(range(0, n, hits[0]))
    counts = [int(v) for v in spec.split(",")] if isinstance(spec, str) else list(spec)
    out, start = [], 0
    for i, cnt in enumerate(counts):
        size = n // len(counts) + (1 if i < n % len(counts) else 0)
        if size < cnt:
            raise ValueError("section too small")
        frac = 1 if cnt <= 1 else (size - 1) / (cnt - 1)
        out += [start + round(c) for c in _strided(cnt, frac)]
        start += size
    return sorted(set(out))


def _strided(cnt, frac):
    cur = 0.0
    for _ in range(cnt):       # accumulated, not multiplied: the rounding of `cur` is the reference's (respace.py:53-57)
        yield cur
        cur += frac


def _base_betas(name, n):
    """gaussian_diffusion.py:26-73."""
    if name == "linear":
        scale = 1000 / n
        return np.linspace(scale * 0.0001, scale * 0.02, n, dtype=np.float64)
    if name == "cosine":
        import math
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        return np.array([min(1 - f((i + 1) / n) / f(i / n), 0.999) for i in range(n)])
    raise NotImplementedError(name)


class GuidedTables:
    """SpacedDiffusion's tables (respace.py:63-99 on gaussian_diffusion.py:119-180); defaults = configs/imagenet.yml."""

    def __init__(self, n=1000, noise_schedule="linear", timestep_respacing="", rescale_timesteps=True):
        ac_base = np.cumprod(1.0 - _base_betas(noise_schedule, n), axis=0)
        self.timestep_map = _kept_steps(n, timestep_respacing)
        new_betas, last = [], 1.0
        for i in self.timestep_map:                                                  # respace.py:76-84
            new_betas.append(1 - ac_base[i] / last)
            last = ac_base[i]
        b = np.array(new_betas, dtype=np.float64)
        self.betas = b
        alphas = 1.0 - b
        self.ac = np.cumprod(alphas, axis=0)
        self.ac_prev = np.append(1.0, self.ac[:-1])
        self.sqrt_recip_ac = np.sqrt(1.0 / self.ac)
        self.sqrt_recipm1_ac = np.sqrt(1.0 / self.ac - 1)
        self.post_var = b * (1.0 - self.ac_prev) / (1.0 - self.ac)
        self.post_logvar_clipped = np.log(np.append(self.post_var[1], self.post_var[1:]))
        self.c1 = b * np.sqrt(self.ac_prev) / (1.0 - self.ac)
        self.c2 = (1.0 - self.ac_prev) * np.sqrt(alphas) / (1.0 - self.ac)
        self.n = n
        self.rescale = rescale_timesteps


def guided_p_sample(unet, tab, x, i, z):
    """One reverse step at integer timestep i for the whole batch."""
    B = x.shape[0]
    t = torch.full((B,), i, dtype=torch.long)
    ts = torch.tensor(tab.timestep_map, dtype=torch.long)[t]                        # respace.py:131-136
    out = unet(x, ts.float() * (1000.0 / tab.n) if tab.rescale else ts.float())
    eps, var = torch.split(out, 3, dim=1)                                           # L272
    min_log = _extract(tab.post_logvar_clipped, t, x.shape)
    max_log = _extract(np.log(tab.betas), t, x.shape)
    frac = (var + 1) / 2
    logvar = frac * max_log + (1 - frac) * min_log                                  # L277-284
    x0 = (_extract(tab.sqrt_recip_ac, t, x.shape) * x - _extract(tab.sqrt_recipm1_ac, t, x.shape) * eps).clamp(-1, 1)
    mean = _extract(tab.c1, t, x.shape) * x0 + _extract(tab.c2, t, x.shape) * x     # L224-227
    nonzero = (t != 0).float().view(-1, 1, 1, 1)
    return mean + nonzero * torch.exp(0.5 * logvar) * z                             # L438-446


def purify_guided(unet, x0, t_levels, init_noise, step_noise, n=1000, **chain):
    tab = GuidedTables(n, **chain)
    betas32 = torch.from_numpy(tab.betas).float()
    a = (1 - betas32).cumprod(dim=0)
    x = x0 * a[t_levels - 1].sqrt() + init_noise * (1.0 - a[t_levels - 1]).sqrt()   # diffpure_guided.py:60-62
    for k, i in enumerate(reversed(range(t_levels))):
        x = guided_p_sample(unet, tab, x, i, step_noise[k])
    return x


def purify_celeba(unet, x0, t_levels, init_noise, step_noise, beta_start=0.0001, beta_end=0.02, n=1000):
    betas64 = np.linspace(beta_start, beta_end, n, dtype=np.float64)                # diffpure_ddpm.py:19-23
    ac64 = np.cumprod(1.0 - betas64, axis=0)
    ac_prev = np.append(1.0, ac64[:-1])
    logvar = np.log(np.maximum(betas64 * (1.0 - ac_prev) / (1.0 - ac64), 1e-20))    # L93-97 (fixedsmall)
    betas = torch.from_numpy(betas64).float()
    a = (1 - betas).cumprod(dim=0)
    x = x0 * a[t_levels - 1].sqrt() + init_noise * (1.0 - a[t_levels - 1]).sqrt()   # L119-120
    B = x.shape[0]
    for k, i in enumerate(reversed(range(t_levels))):
        t = torch.full((B,), i, dtype=torch.long)
        alphas = 1.0 - betas
        ac = alphas.cumprod(dim=0)
        out = unet(x, t)
        wscore = betas / torch.sqrt(1 - ac)
        ex = lambda v: torch.as_tensor(v, dtype=torch.float)[t].view(-1, 1, 1, 1)   # noqa: E731  (extract, L26-34)
        mean = ex(1 / torch.sqrt(alphas)) * (x - ex(wscore) * out)
        mask = (1 - (t == 0).float()).view(-1, 1, 1, 1)
        x = (mean + mask * torch.exp(0.5 * ex(logvar)) * step_noise[k]).float()
    return x
