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


def guided_tables(t_levels, n=1000):
    """Learned-range p_sample of guided_diffusion for timesteps t_levels-1 ... 0
    (runners/diffpure_guided.py:59-75; gaussian_diffusion.py:119-180,240-334,403-447; respace.py:71-99 with
    timestep_respacing '1000', i.e. betas re-derived from the cumulative products; rescale_timesteps -> float t).
    coef row: [sqrt(1/ac), sqrt(1/ac - 1), post_mean_coef1, post_mean_coef2, log beta, post_logvar_clipped, t != 0, 0]."""
    base = _linear_betas64(n)
    ac_base = np.cumprod(1.0 - base, axis=0)
    betas, last = [], 1.0
    for a in ac_base:                                          # SpacedDiffusion.__init__, respace.py:76-84
        betas.append(1 - a / last)
        last = a
    betas = np.array(betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    post_logvar = np.log(np.append(post_var[1], post_var[1:]))
    c1 = betas * np.sqrt(ac_prev) / (1.0 - ac)
    c2 = (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)
    idx = np.arange(t_levels - 1, -1, -1)
    coef = np.stack([np.sqrt(1.0 / ac)[idx], np.sqrt(1.0 / ac - 1)[idx], c1[idx], c2[idx], np.log(betas)[idx],
                     post_logvar[idx], (idx != 0).astype(np.float64), np.zeros(len(idx))], 1).astype(np.float32)
    cond = idx.astype(np.float32) * np.float32(1000.0 / n)     # _WrappedModel, respace.py:131-136
    # forward diffusion uses the fp32 copy of the betas (diffpure_guided.py:39,61-62)
    a32 = (1 - torch.from_numpy(betas).float()).cumprod(dim=0)
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
