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
