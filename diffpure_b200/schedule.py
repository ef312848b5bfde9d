"""Per-step scalar tables of the purification loops, computed on the host with the reference's fp32 op order.

VP-SDE (runners/diffpure_sde.py:86-147,228-239 + torchsde fixed-step Euler):
    t'_0 = fp32(1 - t*/1000), t'_{k+1} = min(t'_k + 1e-3, fp32(1 - 1e-5)) in fp32, h_k = t'_{k+1} - t'_k
    s_k = 1 - t'_k;  beta = 0.1 + 19.9 s_k;  sigma = sqrt(1 - exp(-9.95 s_k^2 - 0.1 s_k))
    x <- x + (beta/2 x - beta/sigma eps) h + sqrt(beta) sqrt(h) z  =  c0 x + c1 eps + c2 z
The device loop consumes (cond_k, c0, c1, c2) per step; nothing is recomputed per step on the device.
"""
import numpy as np
import torch

BETA_MIN, BETA_MAX, N_SCALES = 0.1, 20.0, 1000


def vpsde_time_grid(t_star, dt=1e-3):
    t0, t1 = 1 - t_star * 1. / 1000, 1 - 1e-5
    ts = torch.linspace(t0, t1, 2)
    grid = [ts[0]]
    while grid[-1] < ts[-1]:
        grid.append(torch.minimum(grid[-1] + dt, ts[-1]))
    return torch.stack(grid)


def vpsde_tables(t_star, score_type="score_sde"):
    """Returns (cond [steps] fp32, coef [steps, 3] fp32) for `steps` Euler-Maruyama steps."""
    grid = vpsde_time_grid(t_star)
    t, h = grid[:-1], grid[1:] - grid[:-1]
    s = 1 - t
    beta = BETA_MIN + s * (BETA_MAX - BETA_MIN)
    if score_type == "score_sde":
        cond = s * 999                                                        # models/utils.py:149
        log_mean_coeff = -0.25 * s ** 2 * (BETA_MAX - BETA_MIN) - 0.5 * s * BETA_MIN
        sigma = torch.sqrt(1. - torch.exp(2. * log_mean_coeff))               # sde_lib.py:150-152
    elif score_type == "guided_diffusion":
        cond = (s.float() * N_SCALES).long().float()                           # diffpure_sde.py:82-84,106
        sigma = torch.sqrt(1. - torch.exp(-0.5 * (BETA_MAX - BETA_MIN) * s ** 2 - BETA_MIN * s))
    else:
        raise NotImplementedError(f"Unknown score type in RevVPSDE: {score_type}!")
    c0 = 1 + 0.5 * beta * h
    c1 = -(beta / sigma) * h
    c2 = torch.sqrt(beta) * torch.sqrt(h)
    coef = torch.stack([c0, c1, c2], 1)
    return cond.float().numpy().astype(np.float32), coef.float().numpy().astype(np.float32)


def vpsde_forward_scales(t_level):
    """sqrt(a[t-1]), sqrt(1 - a[t-1]) with a = cumprod(1 - linspace(b0/N, b1/N, N)) in fp32 (diffpure_sde.py:222-223)."""
    betas = torch.linspace(BETA_MIN / N_SCALES, BETA_MAX / N_SCALES, N_SCALES).float()
    a = (1 - betas).cumprod(dim=0)
    return float(a[t_level - 1].sqrt()), float((1.0 - a[t_level - 1]).sqrt())


# ---------------------------------------------------------------------------------------------------------
# DDPM ancestral chains
# ---------------------------------------------------------------------------------------------------------
def _linear_betas64(n, beta_start=None, beta_end=None):
    if beta_start is None:                                     # guided_diffusion get_named_beta_schedule('linear')
        scale = 1000 / n
        beta_start, beta_end = scale * 0.0001, scale * 0.02
    return np.linspace(beta_start, beta_end, n, dtype=np.float64)


def spaced_timesteps(n, section_counts):
    """Sorted base-chain timesteps kept by `timestep_respacing` (respace.py:7-60): the base chain of n steps is cut into
    len(section_counts) near-equal sections and section i keeps section_counts[i] evenly strided steps;
    'ddimK' = the unique integer stride that yields exactly K steps."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, n):
                if len(range(0, n, stride)) == want:
                    return list(range(0, n, stride))
            raise ValueError(f"cannot create exactly {n} steps with an integer stride")
        section_counts = [int(c) for c in section_counts.split(",")]
    q, r = divmod(n, len(section_counts))
    kept, first = set(), 0
    for i, count in enumerate(section_counts):
        size = q + (1 if i < r else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.add(first + round(pos))
            pos += stride
        first += size
    return sorted(kept)


def named_betas64(noise_schedule, n):
    """get_named_beta_schedule (gaussian_diffusion.py:26-73): 'linear' (Ho et al., rescaled to n steps) or 'cosine'
    (betas from alpha_bar(t) = cos^2((t + 0.008) / 1.008 * pi / 2), capped at 0.999)."""
    if noise_schedule == "linear":
        return _linear_betas64(n)
    if noise_schedule == "cosine":
        import math
        bar = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        return np.array([min(1 - bar((i + 1) / n) / bar(i / n), 0.999) for i in range(n)], dtype=np.float64)
    raise NotImplementedError(f"unknown beta schedule: {noise_schedule}")


class GuidedChain:
    """The float64 tables of the reference's SpacedDiffusion (respace.py:63-99 over gaussian_diffusion.py:119-180) for
    (diffusion_steps, noise_schedule, timestep_respacing, rescale_timesteps); '' or None respacing = the full chain.
    Held to the reference's own create_gaussian_diffusion by tests/golden/guided_schedules.npz."""

    def __init__(self, n=1000, noise_schedule="linear", timestep_respacing="", rescale_timesteps=True):
        ac_base = np.cumprod(1.0 - named_betas64(noise_schedule, n), axis=0)
        keep = spaced_timesteps(n, timestep_respacing if timestep_respacing else [n])
        betas, last = [], 1.0
        for i in keep:                                         # SpacedDiffusion.__init__, respace.py:76-84
            betas.append(1 - ac_base[i] / last)
            last = ac_base[i]
        self.n_base, self.rescale = n, bool(rescale_timesteps)
        self.timestep_map = np.array(keep, dtype=np.int64)
        self.betas = betas = np.array(betas, dtype=np.float64)
        alphas = 1.0 - betas
        self.ac = ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        self.sqrt_recip_ac, self.sqrt_recipm1_ac = np.sqrt(1.0 / ac), np.sqrt(1.0 / ac - 1)
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        self.post_logvar_clipped = np.log(np.append(post_var[1], post_var[1:]))
        self.c1 = betas * np.sqrt(ac_prev) / (1.0 - ac)
        self.c2 = (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)
        self.num_timesteps = len(betas)

    def model_timesteps(self, idx):
        """What the UNet is conditioned on at chain index idx: _WrappedModel.__call__ (respace.py:127-136)."""
        ts = self.timestep_map[idx].astype(np.float32)
        return ts * np.float32(1000.0 / self.n_base) if self.rescale else ts


def guided_tables(t_levels, n=1000, noise_schedule="linear", timestep_respacing="", rescale_timesteps=True):
    """Learned-range p_sample of guided_diffusion for chain indices t_levels-1 ... 0
    (runners/diffpure_guided.py:59-75; gaussian_diffusion.py:119-180,240-334,403-447; respace.py:63-136: betas re-derived
    from the kept cumulative products, the model conditioned on the mapped -- and, by default, rescaled -- timestep).
    coef row: [sqrt(1/ac), sqrt(1/ac - 1), post_mean_coef1, post_mean_coef2, log beta, post_logvar_clipped, t != 0, 0]."""
    ch = GuidedChain(n, noise_schedule, timestep_respacing, rescale_timesteps)
    if not 0 < t_levels <= ch.num_timesteps:
        raise ValueError(f"t = {t_levels} outside the {ch.num_timesteps}-step chain")
    idx = np.arange(t_levels - 1, -1, -1)
    coef = np.stack([ch.sqrt_recip_ac[idx], ch.sqrt_recipm1_ac[idx], ch.c1[idx], ch.c2[idx], np.log(ch.betas)[idx],
                     ch.post_logvar_clipped[idx], (idx != 0).astype(np.float64), np.zeros(len(idx))], 1).astype(np.float32)
    cond = ch.model_timesteps(idx)
    # forward diffusion uses the fp32 copy of the betas (diffpure_guided.py:39,61-62)
    a32 = (1 - torch.from_numpy(ch.betas).float()).cumprod(dim=0)
    return cond, coef, float(a32[t_levels - 1].sqrt()), float((1.0 - a32[t_levels - 1]).sqrt())


def ddpm_tables(t_levels, beta_start=0.0001, beta_end=0.02, n=1000, var_type="fixedsmall"):
    """Fixed-variance DDPM step of the CelebA-HQ runner for timesteps t_levels-1 ... 0
    (runners/diffpure_ddpm.py:19-23,37-54,80-97,116-129): x <- c0 x + c1 eps + c2 z with
    c0 = 1/sqrt(alpha_t), c1 = -c0 beta_t/sqrt(1-ac_t) (fp32 tables), c2 = [t != 0] exp(logvar_t / 2)."""
    betas64 = np.linspace(beta_start, beta_end, n, dtype=np.float64)
    alphas64 = 1.0 - betas64
    ac64 = np.cumprod(alphas64, axis=0)
    ac_prev = np.append(1.0, ac64[:-1])
    post_var = betas64 * (1.0 - ac_prev) / (1.0 - ac64)
    if var_type == "fixedlarge":
        logvar = np.log(np.append(post_var[1], betas64[1:]))
    elif var_type == "fixedsmall":
        logvar = np.log(np.maximum(post_var, 1e-20))
    else:
        raise ValueError(var_type)
    betas = torch.from_numpy(betas64).float()                  # L85: fp32 betas drive the step arithmetic
    alphas = 1.0 - betas
    ac = alphas.cumprod(dim=0)
    wscore = betas / torch.sqrt(1 - ac)
    inv = 1 / torch.sqrt(alphas)
    idx = np.arange(t_levels - 1, -1, -1)
    tidx = torch.from_numpy(idx.copy())
    c0 = inv[tidx]
    c1 = -(inv[tidx] * wscore[tidx])
    c2 = torch.exp(0.5 * torch.tensor(logvar, dtype=torch.float)[tidx]) * torch.from_numpy((idx != 0).astype(np.float32))
    coef = torch.stack([c0, c1, c2], 1).numpy().astype(np.float32)
    cond = idx.astype(np.float32)
    return cond, coef, float(ac[t_levels - 1].sqrt()), float((1.0 - ac[t_levels - 1]).sqrt())


# ---------------------------------------------------------------------------------------------------------
# Sibling runners on the same engine
# ---------------------------------------------------------------------------------------------------------
def _sigma(s, score_type):
    if score_type == "score_sde":
        return torch.sqrt(1. - torch.exp(2. * (-0.25 * s ** 2 * (BETA_MAX - BETA_MIN) - 0.5 * s * BETA_MIN)))
    if score_type == "guided_diffusion":
        return torch.sqrt(1. - torch.exp(-0.5 * (BETA_MAX - BETA_MIN) * s ** 2 - BETA_MIN * s))
    raise NotImplementedError(f"Unknown score type in RevVPSDE: {score_type}!")


def _cond(s, score_type):
    return (s * 999) if score_type == "score_sde" else (s.float() * N_SCALES).long().float()


def vpode_tables(t_star, step_size=1e-3, score_type="score_sde"):
    """Probability-flow ODE with torchdiffeq's fixed-grid Euler (runners/diffpure_ode.py:90-131,219-238):
    time runs DOWN from t*/1000 to 1e-5 on the grid t_i = t0 - i*step (last point forced to t1);
    dx/dt = -beta/2 x + beta/(2 sigma) eps, so x <- (1 - beta dt/2) x + beta dt/(2 sigma) eps with dt = t_{i+1} - t_i < 0."""
    t0, t1 = t_star * 1. / 1000, 1e-5
    ts = torch.linspace(t0, t1, 2)
    niters = int(torch.ceil((ts[0] - ts[1]) / step_size + 1).item())          # torchdiffeq _grid_constructor_from_step_size
    grid = ts[0] - torch.arange(0, niters, dtype=ts.dtype) * step_size
    grid[-1] = ts[1]
    s, dt = grid[:-1], grid[1:] - grid[:-1]
    beta = BETA_MIN + s * (BETA_MAX - BETA_MIN)
    sigma = _sigma(s, score_type)
    coef = torch.stack([1 - 0.5 * beta * dt, 0.5 * (beta / sigma) * dt, torch.zeros_like(s)], 1)
    return _cond(s, score_type).float().numpy().astype(np.float32), coef.float().numpy().astype(np.float32)


def ldsde_tables(t_star, sigma2=1e-3, lambda_ld=1e-2, eta=5.0, score_type="score_sde", dt=1e-2):
    """Langevin-dynamics SDE baseline (runners/diffpure_ldsde.py:92-148,196-200,229-243): torchsde Euler with dt = 1e-2 on
    [1 - t*/1000, 1 - 1e-5]; the score is always evaluated at t = 1e-2;
    f = -lambda/2 (eps/sigma + (x - x_init)/sigma2), g = sqrt(lambda) eta  ->  4 coefficients (x, eps, z, x_init)."""
    t0, t1 = 1 - t_star * 1. / 1000, 1 - 1e-5
    ts = torch.linspace(t0, t1, 2)
    grid = [ts[0]]
    while grid[-1] < ts[-1]:
        grid.append(torch.minimum(grid[-1] + dt, ts[-1]))
    grid = torch.stack(grid)
    h = grid[1:] - grid[:-1]
    tfix = torch.zeros_like(h) + 1e-2
    sigma = _sigma(tfix, score_type)
    c0 = 1 - 0.5 * lambda_ld * h / sigma2
    c1 = -0.5 * lambda_ld * h / sigma
    c2 = float(np.sqrt(lambda_ld) * eta) * torch.sqrt(h)
    c3 = 0.5 * lambda_ld * h / sigma2
    coef = torch.stack([c0, c1, c2, c3], 1)
    return _cond(tfix, score_type).float().numpy().astype(np.float32), coef.float().numpy().astype(np.float32)
