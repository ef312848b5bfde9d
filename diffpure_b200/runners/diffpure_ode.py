"""OdeGuidedDiffusion / VPODE with the reference's API (runners/diffpure_ode.py:50-249), backed by the sm_100a engine.

The reference integrates the probability-flow ODE with torchdiffeq's fixed-grid Euler (`method='euler'`,
`options=dict(step_size=args.step_size)`, L226-238): a noise-free linear update per step, i.e. `DP_UPDATE_LINEAR`
with c2 = 0 on the same UNet programs. (The reference reads an undefined `args.fix_rand`, L202; treated as False
when absent.)"""
import os
import random

import numpy as np
import torch

from .. import schedule
from .diffpure_sde import _extract_into_tensor, build_score_model


class VPODE(torch.nn.Module):
    """dx/dt of the probability-flow ODE (L50-131); `forward(t, states)` follows the torchdiffeq protocol."""

    def __init__(self, model, score_type='guided_diffusion', beta_min=0.1, beta_max=20, N=1000,
                 img_shape=(3, 256, 256), model_kwargs=None):
        super().__init__()
        self.model = model
        self.score_type = score_type
        self.model_kwargs = model_kwargs
        self.img_shape = img_shape
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alphas_cumprod_cont = lambda t: torch.exp(-0.5 * (beta_max - beta_min) * t ** 2 - beta_min * t)
        self.sqrt_1m_alphas_cumprod_neg_recip_cont = lambda t: -1. / torch.sqrt(1. - self.alphas_cumprod_cont(t))

    def _scale_timesteps(self, t):
        assert torch.all(t <= 1) and torch.all(t >= 0), f't has to be in [0, 1], but get {t} with shape {t.shape}'
        return (t.float() * self.N).long()

    def ode_fn(self, t, x):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        drift, diffusion = -0.5 * beta_t[:, None] * x, torch.sqrt(beta_t)
        assert x.ndim == 2 and np.prod(self.img_shape) == x.shape[1], x.shape
        x_img = x.view(-1, *self.img_shape)
        if self.score_type == 'guided_diffusion':
            out = self.model(x_img, self._scale_timesteps(t))
            out, _ = torch.split(out, self.img_shape[0], dim=1)
            score = _extract_into_tensor(self.sqrt_1m_alphas_cumprod_neg_recip_cont, t, x.shape) * out.reshape(x.shape[0], -1)
        elif self.score_type == 'score_sde':
            out = self.model(x_img, t * 999)
            std = torch.sqrt(1. - torch.exp(2. * (-0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0)))
            score = (-out / std[:, None, None, None]).reshape(x.shape[0], -1)
        else:
            raise NotImplementedError(f'Unknown score type in RevVPSDE: {self.score_type}!')
        return drift - 0.5 * diffusion[:, None] ** 2 * score

    def forward(self, t, states):
        x = states[0]
        t = t.expand(x.shape[0])
        dx_dt = self.ode_fn(t, x)
        assert dx_dt.shape == x.shape
        return dx_dt,


class OdeGuidedDiffusion(torch.nn.Module):
    def __init__(self, args, config, device=None, state_dict=None):
        super().__init__()
        self.args = args
        self.config = config
        if device is None:
            device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        model, img_shape = build_score_model(config, state_dict)
        self.model = model.eval()
        self.vpode = VPODE(model=model, score_type=args.score_type, img_shape=img_shape, model_kwargs=None)
        self.betas = self.vpode.discrete_betas.float().to(self.device)
        self.atol, self.rtol = 1e-3, 1e-3
        self.method = 'euler'
        self.sample_offset = 0
        print(f'method: {self.method}, atol: {self.atol}, rtol: {self.rtol}, step_size: {self.args.step_size}')

    def image_editing_sample(self, img, bs_id=0, tag=None, init_noise=None):
        assert isinstance(img, torch.Tensor)
        if torch.is_grad_enabled() and img.requires_grad:
            raise NotImplementedError("diffpure_b200: backward through the ODE loop (odeint_adjoint) is not implemented")
        batch_size = img.shape[0]
        if tag is None:
            tag = 'rnd' + str(random.randint(0, 10000))
        out_dir = os.path.join(self.args.log_dir, 'bs' + str(bs_id) + '_' + tag)
        assert img.ndim == 4, img.ndim
        dev = self.device if self.device.type == "cuda" else img.device
        x0 = img.to(dev)
        save = bs_id < 2 and getattr(self.args, "save_images", True)
        if save:
            import torchvision.utils as tvu
            os.makedirs(out_dir, exist_ok=True)
            tvu.save_image((x0 + 1) * 0.5, os.path.join(out_dir, 'original_input.png'))
        eng = self.model.engine_for(batch_size, dev)
        cond, coef = schedule.vpode_tables(self.args.t, self.args.step_size, self.args.score_type)
        sx, se = schedule.vpsde_forward_scales(self.args.t)
        xs = []
        for it in range(self.args.sample_step):
            if getattr(self.args, "fix_rand", False):                      # L202-207
                noise_fixed = torch.FloatTensor(1, *x0.shape[1:]).normal_(
                    0, 1, generator=torch.manual_seed(self.args.seed)).to(dev)
                e = noise_fixed.repeat(x0.shape[0], 1, 1, 1)
            else:
                e = torch.randn_like(x0) if init_noise is None else init_noise.to(dev)
            x0 = eng.purify(x0, cond, coef, sx, se, init_noise=e, sample_offset=self.sample_offset)
            if save:
                import torchvision.utils as tvu
                torch.save(x0, os.path.join(out_dir, f'samples_{it}.pth'))
                tvu.save_image((x0 + 1) * 0.5, os.path.join(out_dir, f'samples_{it}.png'))
            xs.append(x0)
        return torch.cat(xs, dim=0)
