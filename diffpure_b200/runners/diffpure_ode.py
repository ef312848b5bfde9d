"""OdeGuidedDiffusion / VPODE with the reference's API (runners/diffpure_ode.py:50-249), backed by the sm_100a engine.

The reference integrates the probability-flow ODE with torchdiffeq's fixed-grid Euler (`method='euler'`,
`options=dict(step_size=args.step_size)`, L226-238): a noise-free linear update per step, i.e. `DP_UPDATE_LINEAR`
with c2 = 0 on the same UNet programs. (The reference reads an undefined `args.fix_rand`, L202; treated as False
when absent.)"""
import torch

from .. import lib as _lib
from .. import schedule
from ._common import PurifyRunner, PurifyWithGrad, VPScore
from .diffpure_sde import build_score_model


class VPODE(VPScore):
    """dx/dt of the probability-flow ODE (L50-131); `forward(t, states)` follows the torchdiffeq protocol."""

    def __init__(self, model, score_type='guided_diffusion', beta_min=0.1, beta_max=20, N=1000,
                 img_shape=(3, 256, 256), model_kwargs=None):
        super().__init__(model, score_type, beta_min, beta_max, N, img_shape, model_kwargs)

    def ode_fn(self, t, x):
        """-beta_t/2 x - beta_t/2 score(x, t) at forward time t (B,)."""
        beta_t = self.beta(t)
        return -0.5 * beta_t[:, None] * x - 0.5 * beta_t[:, None] * self.score(t, x)

    def forward(self, t, states):
        x = states[0]
        dx_dt = self.ode_fn(t.expand(x.shape[0]), x)
        assert dx_dt.shape == x.shape
        return dx_dt,


class OdeGuidedDiffusion(PurifyRunner):
    def __init__(self, args, config, device=None, state_dict=None):
        super().__init__()
        self._setup(args, config, device)
        model, img_shape = build_score_model(config, state_dict)
        self.model = model.eval()
        self.vpode = VPODE(model=model, score_type=args.score_type, img_shape=img_shape, model_kwargs=None)
        self.betas = self.vpode.discrete_betas.float().to(self.device)
        self.atol, self.rtol = 1e-3, 1e-3
        self.method = 'euler'
        print(f'method: {self.method}, atol: {self.atol}, rtol: {self.rtol}, step_size: {self.args.step_size}')

    def image_editing_sample(self, img, bs_id=0, tag=None, init_noise=None):
        x0, dev, dump = self._open(img, bs_id, tag)
        eng = self.model.engine_for(x0.shape[0], dev)
        cond, coef = schedule.vpode_tables(self.args.t, self.args.step_size, self.args.score_type)
        sx, se = schedule.vpsde_forward_scales(self.args.t)

        def one_pass(it, x):
            if getattr(self.args, "fix_rand", False):                      # one shared draw for the whole batch, L202-207
                fixed = torch.FloatTensor(1, *x.shape[1:]).normal_(0, 1, generator=torch.manual_seed(self.args.seed))
                e = fixed.to(dev).repeat(x.shape[0], 1, 1, 1)
            else:
                e = self._init_noise(x, init_noise, dev)
            if dump.on:
                dump.image(f'init_{it}.png', (x * sx + e * se).detach())
            if self._wants_grad(x):      # odeint_adjoint in the reference (L230-238): here the discrete Euler loop's gradient
                return PurifyWithGrad.apply(x, None, self.model, cond, coef, sx, se, e, None, 0, self.sample_offset,
                                            _lib.DP_UPDATE_LINEAR)
            return eng.purify(x, cond, coef, sx, se, init_noise=e, sample_offset=self.sample_offset, **self._fuse_kw)

        return self._passes(x0, dump, one_pass)
