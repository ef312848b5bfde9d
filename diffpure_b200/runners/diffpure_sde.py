"""RevVPSDE / RevGuidedDiffusion with the reference's API (runners/diffpure_sde.py:50-247), backed by the
sm_100a engine. `eval_sde_adv.SDE_Adv_Model` can import this module in place of the reference's.

Differences, all stated in DESIGN.md:
  * the Euler-Maruyama loop (torchsde in the reference, L233-239) runs on the device inside the engine
    (`dp_purify`): UNet evaluation + fused update per step, no host synchronisation inside the loop;
  * Brownian increments come from the engine's counter-based generator keyed by (seed, global sample index,
    step, pixel) -- the seed is drawn from NumPy's global RNG as torchsde's BrownianInterval does when no
    entropy is given -- or are injected (`step_noise=`) for parity tests;
  * gradients (white-box attacks, eval_sde_adv.py:126-128): the reference differentiates with torchsde's continuous
    adjoint; here `image_editing_sample` is a torch.autograd.Function whose backward is the exact gradient of the
    discrete Euler-Maruyama loop the engine runs (discretise-then-differentiate): the forward pass records the states
    x_k, the backward pass replays them through the engine's UNet vector-Jacobian program (`dp_unet_vjp`),
    lambda_k = c0_k lambda_{k+1} + J_k^T (c1_k lambda_{k+1}). Available for the DDPM++ network (CIFAR-10).
"""
import numpy as np
import torch

from .. import lib as _lib
from .. import schedule
from ..model import ScoreModel
from ._common import PurifyRunner, PurifyWithGrad, VPScore, _extract_into_tensor  # noqa: F401  (re-exported like the reference module)


class RevVPSDE(VPScore):
    """The torchsde "SDE object" of the reference (L50-147): f/g of the time-reversed VP-SDE on flattened states, usable
    with an external torchsde.sdeint; the engine-backed `model` evaluates the score network."""

    def __init__(self, model, score_type='guided_diffusion', beta_min=0.1, beta_max=20, N=1000,
                 img_shape=(3, 256, 256), model_kwargs=None):
        super().__init__(model, score_type, beta_min, beta_max, N, img_shape, model_kwargs)
        self.noise_type = "diagonal"
        self.sde_type = "ito"

    def vpsde_fn(self, t, x):
        """Forward VP-SDE: drift -beta_t/2 x, diffusion sqrt(beta_t)."""
        beta_t = self.beta(t)
        return -0.5 * beta_t[:, None] * x, torch.sqrt(beta_t)

    def rvpsde_fn(self, t, x, return_type='drift'):
        """Reverse-time SDE at forward time t: drift f - g^2 score, same diffusion (L98-129)."""
        drift, diffusion = self.vpsde_fn(t, x)
        if return_type != 'drift':
            return diffusion
        return drift - diffusion[:, None] ** 2 * self.score(t, x)

    def f(self, t, x):
        """torchsde drift at reversed time t (scalar tensor): sign-flipped reverse drift at 1 - t."""
        drift = self.rvpsde_fn(1 - t.expand(x.shape[0]), x, return_type='drift')
        assert drift.shape == x.shape
        return -drift

    def g(self, t, x):
        diffusion = self.rvpsde_fn(1 - t.expand(x.shape[0]), x, return_type='diffusion')
        assert diffusion.shape == (x.shape[0],)
        return diffusion[:, None].expand(x.shape)


def _load_score_sde_state(path, device="cpu"):
    """checkpoint_8.pth layout (runners/diffpure_sde.py:42-47,178-182): EMA shadow params replace the weights
    (score_sde/models/ema.py:61-72 copies them in parameter order)."""
    state = torch.load(path, map_location=device)
    sd = {k.replace("module.", "", 1) if k.startswith("module.") else k: v for k, v in state["model"].items()}
    names = [k for k in sd if k != "sigmas"]
    shadow = state["ema"]["shadow_params"]
    assert len(shadow) == len(names), "EMA shadow parameter count does not match the model"
    for k, v in zip(names, shadow):
        sd[k] = v
    return sd


def build_score_model(config, state_dict=None):
    """Dataset -> engine-backed score model, as runners/diffpure_sde.py:160-187 (shared by the SDE / ODE / LDSDE
    runners, which load the same checkpoints)."""
    if config.data.dataset == 'ImageNet':
        from .. import lowering_adm
        cfg = lowering_adm.cfg_from_reference(config)
        img_shape = (3, cfg.image_size, cfg.image_size)
        if state_dict is None:
            state_dict = torch.load('pretrained/guided_diffusion/256x256_diffusion_uncond.pt', map_location='cpu')
        model = ScoreModel("adm", cfg, state_dict, lowering_adm.lower, out_channels=6,
                           lower_vjp_fn=lowering_adm.lower_vjp)
    elif config.data.dataset == 'CIFAR10':
        from .. import lowering_ncsnpp
        cfg = lowering_ncsnpp.cfg_from_reference(config)
        img_shape = (cfg.num_channels, cfg.image_size, cfg.image_size)
        if state_dict is None:
            state_dict = _load_score_sde_state('pretrained/score_sde/checkpoint_8.pth')
        model = ScoreModel("ncsnpp", cfg, state_dict, lowering_ncsnpp.lower, out_channels=cfg.num_channels,
                           lower_vjp_fn=lowering_ncsnpp.lower_vjp)
    else:
        raise NotImplementedError(f'Unknown dataset {config.data.dataset}!')
    return model, img_shape


class RevGuidedDiffusion(PurifyRunner):

    def __init__(self, args, config, device=None, state_dict=None):
        """Same arguments as the reference (L151); `state_dict` optionally supplies the UNet weights
        (reference parameter names) instead of the pretrained checkpoint files."""
        super().__init__()
        self._setup(args, config, device)
        model, img_shape = build_score_model(config, state_dict)
        self.model = model.eval()
        self.img_shape = img_shape
        self.rev_vpsde = RevVPSDE(model=model, score_type=args.score_type, img_shape=img_shape, model_kwargs=None)
        self.betas = self.rev_vpsde.discrete_betas.float().to(self.device)
        self._tables = {}
        print(f't: {args.t}, rand_t: {args.rand_t}, t_delta: {args.t_delta}')
        print(f'use_bm: {args.use_bm}')

    def _tables_for(self, t_star):
        key = (int(t_star), self.args.score_type)
        if key not in self._tables:
            self._tables[key] = schedule.vpsde_tables(t_star, self.args.score_type)
        return self._tables[key]

    def image_editing_sample(self, img, bs_id=0, tag=None, init_noise=None, step_noise=None, seed=None):
        """Reference signature (L197) plus optional injected noise for parity tests.
        img: [B,3,H,W] in [-1,1]. Returns the purified batch(es), `sample_step` of them concatenated."""
        x0, dev, dump = self._open(img, bs_id, tag)
        eng = self.model.engine_for(x0.shape[0], dev)
        cond, coef = self._tables_for(self.args.t)      # the reverse grid always spans t* (L228-231) ...

        def one_pass(it, x):
            e = self._init_noise(x, init_noise, dev)                                     # L217
            level = self.args.t
            if self.args.rand_t:                         # ... only the forward-diffusion level is jittered (L219-223)
                level = self.args.t + np.random.randint(-self.args.t_delta, self.args.t_delta)
                print(f'total_noise_levels: {level}')
            sx, se = schedule.vpsde_forward_scales(level)
            s = self._call_seed(seed, it)
            if self._wants_grad(x):
                dump.image(f'init_{it}.png', (x * sx + e * se).detach())
                return PurifyWithGrad.apply(x, None, self.model, cond, coef, sx, se, e, step_noise, s,
                                            self.sample_offset, _lib.DP_UPDATE_LINEAR)
            if dump.on:
                dump.image(f'init_{it}.png', x * sx + e * se)
            return eng.purify(x, cond, coef, sx, se, init_noise=e, step_noise=step_noise, seed=s,
                              sample_offset=self.sample_offset, **self._fuse_kw)       # L228-239

        return self._passes(x0, dump, one_pass)
