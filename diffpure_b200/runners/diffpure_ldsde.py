"""LDGuidedDiffusion / LDSDE with the reference's API (runners/diffpure_ldsde.py:50-252), backed by the sm_100a engine.

Langevin-dynamics baseline: torchsde Euler (dt = 1e-2, L196-200) on f = -lambda/2 (eps/sigma + (x - x_init)/sigma2),
g = sqrt(lambda) eta, the score always evaluated at t = 1e-2 (L94). Linear in (x, eps, z, x_init):
`DP_UPDATE_LINEAR_ANCHORED`, no forward diffusion (x starts at the input, L219)."""
import numpy as np
import torch

from .. import lib as _lib
from .. import schedule
from ._common import PurifyRunner, PurifyWithGrad, VPScore
from .diffpure_sde import build_score_model


class LDSDE(VPScore):
    """The torchsde SDE object of the reference (L50-148)."""

    SCORE_TIME = 1e-2   # the score network is always queried at this forward time

    def __init__(self, model, x_init, score_type='guided_diffusion', beta_min=0.1, beta_max=20, N=1000,
                 img_shape=(3, 256, 256), sigma2=0.001, lambda_ld=0.01, eta=5, model_kwargs=None):
        super().__init__(model, score_type, beta_min, beta_max, N, img_shape, model_kwargs)
        self.x_init = x_init
        self.sigma2, self.eta, self.lambda_ld = sigma2, eta, lambda_ld
        self.noise_type = "diagonal"
        self.sde_type = "ito"
        print(f'sigma2: {self.sigma2}, lambda_ld: {self.lambda_ld}, eta: {self.eta}')

    def ldsde_fn(self, t, x, return_type='drift'):
        if return_type != 'drift':
            g = float(np.sqrt(self.lambda_ld) * self.eta)
            return torch.full((x.shape[0],), g, dtype=torch.float, device=x.device)
        t = torch.zeros_like(t, dtype=torch.float, device=t.device) + self.SCORE_TIME
        pull = (x - self.x_init) / self.sigma2              # towards the input image
        return -0.5 * (-self.score(t, x) + pull) * self.lambda_ld

    def f(self, t, x):
        return self.ldsde_fn(t.expand(x.shape[0]), x, return_type='drift')

    def g(self, t, x):
        return self.ldsde_fn(t.expand(x.shape[0]), x, return_type='diffusion')[:, None].expand(x.shape)


class LDGuidedDiffusion(PurifyRunner):
    def __init__(self, args, config, device=None, state_dict=None):
        super().__init__()
        self._setup(args, config, device)
        model, img_shape = build_score_model(config, state_dict)
        self.model = model.eval()
        self.img_shape = img_shape
        self.args_dict = {'method': 'euler', 'adaptive': False, 'dt': 1e-2}
        print(f'args_dict: {self.args_dict}')

    def image_editing_sample(self, img, bs_id=0, tag=None, step_noise=None, seed=None):
        x0, dev, dump = self._open(img, bs_id, tag)
        a = self.args
        self.ldsde = LDSDE(model=self.model, x_init=x0.view(x0.shape[0], -1), score_type=a.score_type,
                           img_shape=self.img_shape, sigma2=a.sigma2, lambda_ld=a.lambda_ld, eta=a.eta, model_kwargs=None)
        self.betas = self.ldsde.discrete_betas.float().to(dev)
        eng = self.model.engine_for(x0.shape[0], dev)
        cond, coef = schedule.ldsde_tables(a.t, a.sigma2, a.lambda_ld, a.eta, a.score_type, self.args_dict['dt'])
        anchor = x0                                         # L216: x_init is the input image for every pass

        def one_pass(it, x):
            dump.image(f'init_{it}.png', x.detach())
            sd_ = self._call_seed(seed, it)
            if self._wants_grad(x) or self._wants_grad(anchor):
                return PurifyWithGrad.apply(x, anchor, self.model, cond, coef, 1.0, 0.0, torch.zeros_like(x), step_noise,
                                            sd_, self.sample_offset, _lib.DP_UPDATE_LINEAR_ANCHORED)
            return eng.purify(x, cond, coef, 1.0, 0.0, update_kind=_lib.DP_UPDATE_LINEAR_ANCHORED,
                              init_noise=torch.zeros_like(x), step_noise=step_noise, seed=sd_,
                              sample_offset=self.sample_offset, anchor=anchor)

        return self._passes(x0, dump, one_pass)
