"""LDGuidedDiffusion / LDSDE with the reference's API (runners/diffpure_ldsde.py:50-252), backed by the sm_100a engine.

Langevin-dynamics baseline: torchsde Euler (dt = 1e-2, L196-200) on f = -lambda/2 (eps/sigma + (x - x_init)/sigma2),
g = sqrt(lambda) eta, the score always evaluated at t = 1e-2 (L94). Linear in (x, eps, z, x_init):
`DP_UPDATE_LINEAR_ANCHORED`, no forward diffusion (x starts at the input, L219)."""
import os
import random

import numpy as np
import torch

from .. import lib as _lib
from .. import schedule
from .diffpure_sde import _extract_into_tensor, build_score_model


class LDSDE(torch.nn.Module):
    """The torchsde SDE object of the reference (L50-148)."""

    def __init__(self, model, x_init, score_type='guided_diffusion', beta_min=0.1, beta_max=20, N=1000,
                 img_shape=(3, 256, 256), sigma2=0.001, lambda_ld=0.01, eta=5, model_kwargs=None):
        super().__init__()
        self.model, self.x_init = model, x_init
        self.sigma2, self.eta, self.lambda_ld = sigma2, eta, lambda_ld
        self.score_type, self.model_kwargs, self.img_shape = score_type, model_kwargs, img_shape
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas_cumprod_cont = lambda t: torch.exp(-0.5 * (beta_max - beta_min) * t ** 2 - beta_min * t)
        self.sqrt_1m_alphas_cumprod_neg_recip_cont = lambda t: -1. / torch.sqrt(1. - self.alphas_cumprod_cont(t))
        self.noise_type = "diagonal"
        self.sde_type = "ito"
        print(f'sigma2: {self.sigma2}, lambda_ld: {self.lambda_ld}, eta: {self.eta}')

    def _scale_timesteps(self, t):
        assert torch.all(t <= 1) and torch.all(t >= 0), f't has to be in [0, 1], but get {t} with shape {t.shape}'
        return (t.float() * self.N).long()

    def ldsde_fn(self, t, x, return_type='drift'):
        t = torch.zeros_like(t, dtype=torch.float, device=t.device) + 1e-2
        if return_type != 'drift':
            diffusion_coef = np.sqrt(self.lambda_ld) * self.eta
            return torch.tensor([diffusion_coef], dtype=torch.float).expand(x.shape[0]).to(x.device)
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
        return -0.5 * (-score + (x - self.x_init) / self.sigma2) * self.lambda_ld

    def f(self, t, x):
        t = t.expand(x.shape[0])
        return self.ldsde_fn(t, x, return_type='drift')

    def g(self, t, x):
        t = t.expand(x.shape[0])
        return self.ldsde_fn(t, x, return_type='diffusion')[:, None].expand(x.shape)


class LDGuidedDiffusion(torch.nn.Module):
    def __init__(self, args, config, device=None, state_dict=None):
        super().__init__()
        self.args = args
        self.config = config
        if device is None:
            device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        model, img_shape = build_score_model(config, state_dict)
        self.model = model.eval()
        self.img_shape = img_shape
        self.args_dict = {'method': 'euler', 'adaptive': False, 'dt': 1e-2}
        self.sample_offset = 0
        self.last_seed = None
        print(f'args_dict: {self.args_dict}')

    def image_editing_sample(self, img, bs_id=0, tag=None, step_noise=None, seed=None):
        assert isinstance(img, torch.Tensor)
        if torch.is_grad_enabled() and img.requires_grad:
            raise NotImplementedError("diffpure_b200: backward through the LDSDE loop (sdeint_adjoint) is not implemented")
        batch_size = img.shape[0]
        if tag is None:
            tag = 'rnd' + str(random.randint(0, 10000))
        out_dir = os.path.join(self.args.log_dir, 'bs' + str(bs_id) + '_' + tag)
        assert img.ndim == 4, img.ndim
        dev = self.device if self.device.type == "cuda" else img.device
        x0 = img.to(dev)
        self.ldsde = LDSDE(model=self.model, x_init=x0.view(batch_size, -1), score_type=self.args.score_type,
                           img_shape=self.img_shape, sigma2=self.args.sigma2, lambda_ld=self.args.lambda_ld,
                           eta=self.args.eta, model_kwargs=None)
        self.betas = self.ldsde.discrete_betas.float().to(dev)
        save = bs_id < 2 and getattr(self.args, "save_images", True)
        if save:
            import torchvision.utils as tvu
            os.makedirs(out_dir, exist_ok=True)
            tvu.save_image((x0 + 1) * 0.5, os.path.join(out_dir, 'original_input.png'))
        eng = self.model.engine_for(batch_size, dev)
        cond, coef = schedule.ldsde_tables(self.args.t, self.args.sigma2, self.args.lambda_ld, self.args.eta,
                                           self.args.score_type, self.args_dict['dt'])
        anchor = x0                                                        # L216: x_init is the input image for every pass
        xs = []
        for it in range(self.args.sample_step):
            call_seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed) + it
            self.last_seed = call_seed
            x0 = eng.purify(x0, cond, coef, 1.0, 0.0, update_kind=_lib.DP_UPDATE_LINEAR_ANCHORED,
                            init_noise=torch.zeros_like(x0), step_noise=step_noise, seed=call_seed,
                            sample_offset=self.sample_offset, anchor=anchor)
            if save:
                import torchvision.utils as tvu
                torch.save(x0, os.path.join(out_dir, f'samples_{it}.pth'))
                tvu.save_image((x0 + 1) * 0.5, os.path.join(out_dir, f'samples_{it}.png'))
            xs.append(x0)
        return torch.cat(xs, dim=0)
