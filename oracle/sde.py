"""Oracle restatement of the purification loops (CPU, fp32) -- test infrastructure only.

VP-SDE path: runners/diffpure_sde.py:197-247 (forward diffusion L217-223, time grid L228-231) driving
RevVPSDE.f / g (L86-147) through torchsde's fixed-step Ito Euler-Maruyama (`sdeint_adjoint(..., method='euler')`,
third-party, unpinned, absent offline): y <- y + f(t,y) dt + g(t,y) dW with dt = 1e-3, the last step clipped to
ts[-1], dW ~ N(0, dt) [upstream behaviour; independently restated by score_sde/sampling.py:177-187
EulerMaruyamaPredictor with score_sde/sde_lib.py:79-117]. The Brownian increments are injected explicitly
(`step_noise` standard normals z_k, dW_k = sqrt(dt_k) z_k), which is the reference's own `bm` hook (L234-236).
"""
import numpy as np
import torch

BETA_MIN, BETA_MAX, N_SCALES = 0.1, 20.0, 1000


def time_grid(t_star, dt=1e-3):
    """fp32 step grid t'_k of the torchsde loop from t0 = 1 - t*/1000 to t1 = 1 - 1e-5 (diffpure_sde.py:228-231)."""
    t0, t1 = 1 - t_star * 1. / 1000 + 0, 1 - 1e-5
    ts = torch.linspace(t0, t1, 2)                                  # fp32
    grid = [ts[0]]
    while grid[-1] < ts[-1]:
        grid.append(torch.minimum(grid[-1] + dt, ts[-1]))           # fp32 accumulation, last step clipped
    return torch.stack(grid)


def forward_diffuse(x0, noise, t_level):
    """diffpure_sde.py:217-223: x = x0*sqrt(a[t-1]) + e*sqrt(1-a[t-1]), a = cumprod(1-betas) in fp32."""
    betas = torch.linspace(BETA_MIN / N_SCALES, BETA_MAX / N_SCALES, N_SCALES).float()
    a = (1 - betas).cumprod(dim=0)
    return x0 * a[t_level - 1].sqrt() + noise * (1.0 - a[t_level - 1]).sqrt()


def rev_vpsde_f(unet, score_type, t, x_img):
    """RevVPSDE.f (diffpure_sde.py:131-138) for image-shaped x; t is the reversed time t' (scalar fp32 tensor)."""
    B = x_img.shape[0]
    s = 1 - t.expand(B)                                             # forward time, L136
    beta = BETA_MIN + s * (BETA_MAX - BETA_MIN)                     # vpsde_fn L87
    drift = -0.5 * beta[:, None, None, None] * x_img
    diffusion = torch.sqrt(beta)
    if score_type == "score_sde":                                   # L114-120, models/utils.py:144-158
        labels = s * 999
        out = unet(x_img, labels)
        log_mean_coeff = -0.25 * s ** 2 * (BETA_MAX - BETA_MIN) - 0.5 * s * BETA_MIN   # sde_lib.py:150
        std = torch.sqrt(1. - torch.exp(2. * log_mean_coeff))
        score = -out / std[:, None, None, None]
    elif score_type == "guided_diffusion":                          # L101-112
        disc = (s.float() * N_SCALES).long()
        out = unet(x_img, disc)[:, :3]
        ac = torch.exp(-0.5 * (BETA_MAX - BETA_MIN) * s ** 2 - BETA_MIN * s)
        score = (-1. / torch.sqrt(1. - ac))[:, None, None, None] * out
    else:
        raise NotImplementedError(score_type)
    drift = drift - diffusion[:, None, None, None] ** 2 * score     # L125
    return -drift                                                   # L138


def rev_vpsde_g(t, B):
    s = 1 - t.expand(B)
    return torch.sqrt(BETA_MIN + s * (BETA_MAX - BETA_MIN))         # L140-147


def purify_sde(unet, x0, t_star, init_noise, step_noise, score_type="score_sde", t_level=None):
    """image_editing_sample for one sample_step. step_noise: [steps, B, 3, H, W] standard normals."""
    x = forward_diffuse(x0, init_noise, t_star if t_level is None else t_level)
    grid = time_grid(t_star)
    B = x.shape[0]
    for k in range(len(grid) - 1):
        t, t_next = grid[k], grid[k + 1]
        h = t_next - t
        f = rev_vpsde_f(unet, score_type, t, x)
        g = rev_vpsde_g(t, B)[:, None, None, None]
        dW = step_noise[k] * torch.sqrt(h)
        x = x + f * h + g * dW                                       # torchsde Euler (Ito, diagonal noise)
    return x


def num_steps(t_star):
    return len(time_grid(t_star)) - 1


# ---------------------------------------------------------------------------------------------------------
# Sibling loops (same UNet, different integrand): probability-flow ODE and Langevin-dynamics SDE
# ---------------------------------------------------------------------------------------------------------
def _score(unet, score_type, s, x_img):
    """score at forward time s (vector [B]); shared by VPODE.ode_fn (diffpure_ode.py:90-125) and LDSDE.ldsde_fn."""
    if score_type == "score_sde":
        out = unet(x_img, s * 999)
        std = torch.sqrt(1. - torch.exp(2. * (-0.25 * s ** 2 * (BETA_MAX - BETA_MIN) - 0.5 * s * BETA_MIN)))
        return -out / std[:, None, None, None]
    disc = (s.float() * N_SCALES).long()
    out = unet(x_img, disc)[:, :3]
    ac = torch.exp(-0.5 * (BETA_MAX - BETA_MIN) * s ** 2 - BETA_MIN * s)
    return (-1. / torch.sqrt(1. - ac))[:, None, None, None] * out


def vpode_f(unet, score_type, t, x_img):
    """VPODE.ode_fn, diffpure_ode.py:90-125: drift - g^2/2 * score at forward time t."""
    B = x_img.shape[0]
    s = t.expand(B)
    beta = BETA_MIN + s * (BETA_MAX - BETA_MIN)
    return -0.5 * beta[:, None, None, None] * x_img - 0.5 * beta[:, None, None, None] * _score(unet, score_type, s, x_img)


def purify_ode(unet, x0, t_star, init_noise, step_size=1e-3, score_type="score_sde"):
    """OdeGuidedDiffusion.image_editing_sample with method='euler' (diffpure_ode.py:219-241); torchdiffeq's fixed-grid
    Euler (third party, torchdiffeq==0.2.1, absent offline) restated: t_i = t0 - i*step, last point = t1, y += dt*f."""
    x = forward_diffuse(x0, init_noise, t_star)
    t0, t1 = t_star * 1. / 1000, 1e-5
    ts = torch.linspace(t0, t1, 2)
    niters = int(torch.ceil((ts[0] - ts[1]) / step_size + 1).item())
    grid = ts[0] - torch.arange(0, niters, dtype=ts.dtype) * step_size
    grid[-1] = ts[1]
    for i in range(len(grid) - 1):
        x = x + (grid[i + 1] - grid[i]) * vpode_f(unet, score_type, grid[i], x)
    return x


def ldsde_f(unet, score_type, x_img, x_init, sigma2, lambda_ld):
    """LDSDE.ldsde_fn drift, diffpure_ldsde.py:92-131 (score always at t = 1e-2)."""
    B = x_img.shape[0]
    s = torch.zeros(B) + 1e-2
    return -0.5 * (-_score(unet, score_type, s, x_img) + (x_img - x_init) / sigma2) * lambda_ld


def purify_ldsde(unet, x0, t_star, step_noise, sigma2=1e-3, lambda_ld=1e-2, eta=5.0, score_type="score_sde", dt=1e-2):
    """LDGuidedDiffusion.image_editing_sample (diffpure_ldsde.py:205-247): no forward diffusion, torchsde Euler, dt 1e-2."""
    t0, t1 = 1 - t_star * 1. / 1000, 1 - 1e-5
    ts = torch.linspace(t0, t1, 2)
    t = ts[0]
    x = x0
    k = 0
    g = float(np.sqrt(lambda_ld) * eta)
    while t < ts[-1]:
        tn = torch.minimum(t + dt, ts[-1])
        h = tn - t
        x = x + ldsde_f(unet, score_type, x, x0, sigma2, lambda_ld) * h + g * step_noise[k] * torch.sqrt(h)
        t = tn
        k += 1
    return x


def num_steps_ldsde(t_star, dt=1e-2):
    t0, t1 = 1 - t_star * 1. / 1000, 1 - 1e-5
    ts = torch.linspace(t0, t1, 2)
    t, n = ts[0], 0
    while t < ts[-1]:
        t = torch.minimum(t + dt, ts[-1])
        n += 1
    return n
