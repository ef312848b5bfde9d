"""RevVPSDE / RevGuidedDiffusion with the reference's API (runners/diffpure_sde.py:50-247), backed by the
sm_100a engine. `eval_sde_adv.SDE_Adv_Model` can import this module in place of the reference's.

Differences, all stated in DESIGN.md:
  * the Euler-Maruyama loop (torchsde in the reference, L233-239) runs on the device inside the engine
    (`dp_purify`): UNet evaluation + fused update per step, no host synchronisation inside the loop;
  * Brownian increments come from the engine's counter-based generator keyed by (seed, global sample index,
    step, pixel) -- the seed is drawn from NumPy's global RNG as torchsde's BrownianInterval does when no
    entropy is given -- or are injected (`step_noise=`) for parity tests;
  * forward only: differentiating through the loop (torchsde adjoint) is not implemented and raises.
"""
import os
import random

import numpy as np
import torch

from .. import schedule
from ..model import ScoreModel


def _extract_into_tensor(arr_or_func, timesteps, broadcast_shape):
    """runners/diffpure_sde.py:23-39."""
    if callable(arr_or_func):
        res = arr_or_func(timesteps).float()
    else:
        res = arr_or_func.to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


class RevVPSDE(torch.nn.Module):
    """The torchsde "SDE object" of the reference (L50-147): f/g on flattened states, usable with an external
    torchsde.sdeint; the engine-backed `model` evaluates the score network."""

    def __init__(self, model, score_type='guided_diffusion', beta_min=0.1, beta_max=20, N=1000,
                 img_shape=(3, 256, 256), model_kwargs=None):
        super().__init__()
        self.model = model
        self.score_type = score_type
        self.model_kwargs = model_kwargs
        self.img_shape = img_shape

        self.beta_0 = beta_min
        self.beta_1 = beta_max
        self.N = N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - self.alphas_cumprod)

        self.alphas_cumprod_cont = lambda t: torch.exp(-0.5 * (beta_max - beta_min) * t ** 2 - beta_min * t)
        self.sqrt_1m_alphas_cumprod_neg_recip_cont = lambda t: -1. / torch.sqrt(1. - self.alphas_cumprod_cont(t))

        self.noise_type = "diagonal"
        self.sde_type = "ito"

    def _scale_timesteps(self, t):
        assert torch.all(t <= 1) and torch.all(t >= 0), f't has to be in [0, 1], but get {t} with shape {t.shape}'
        return (t.float() * self.N).long()

    def vpsde_fn(self, t, x):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        drift = -0.5 * beta_t[:, None] * x
        diffusion = torch.sqrt(beta_t)
        return drift, diffusion

    def rvpsde_fn(self, t, x, return_type='drift'):
        drift, diffusion = self.vpsde_fn(t, x)
        if return_type != 'drift':
            return diffusion
        assert x.ndim == 2 and np.prod(self.img_shape) == x.shape[1], x.shape
        x_img = x.view(-1, *self.img_shape)
        if self.score_type == 'guided_diffusion':
            disc_steps = self._scale_timesteps(t)
            model_output = self.model(x_img, disc_steps)
            model_output, _ = torch.split(model_output, self.img_shape[0], dim=1)
            model_output = model_output.reshape(x.shape[0], -1)
            score = _extract_into_tensor(self.sqrt_1m_alphas_cumprod_neg_recip_cont, t, x.shape) * model_output
        elif self.score_type == 'score_sde':
            labels = t * 999                                               # score_sde/models/utils.py:149
            out = self.model(x_img, labels)
            log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
            std = torch.sqrt(1. - torch.exp(2. * log_mean_coeff))          # sde_lib.py:149-153
            score = (-out / std[:, None, None, None]).reshape(x.shape[0], -1)
        else:
            raise NotImplementedError(f'Unknown score type in RevVPSDE: {self.score_type}!')
        return drift - diffusion[:, None] ** 2 * score

    def f(self, t, x):
        t = t.expand(x.shape[0])
        drift = self.rvpsde_fn(1 - t, x, return_type='drift')
        assert drift.shape == x.shape
        return -drift

    def g(self, t, x):
        t = t.expand(x.shape[0])
        diffusion = self.rvpsde_fn(1 - t, x, return_type='diffusion')
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
        img_shape = (3, 256, 256)
        model_dir = 'pretrained/guided_diffusion'
        cfg = lowering_adm.cfg_from_reference(config)
        img_shape = (3, cfg.image_size, cfg.image_size)
        if state_dict is None:
            state_dict = torch.load(f'{model_dir}/256x256_diffusion_uncond.pt', map_location='cpu')
        model = ScoreModel("adm", cfg, state_dict, lowering_adm.lower, out_channels=6)
    elif config.data.dataset == 'CIFAR10':
        from .. import lowering_ncsnpp
        model_dir = 'pretrained/score_sde'
        cfg = lowering_ncsnpp.cfg_from_reference(config)
        img_shape = (cfg.num_channels, cfg.image_size, cfg.image_size)
        if state_dict is None:
            state_dict = _load_score_sde_state(f'{model_dir}/checkpoint_8.pth')
        model = ScoreModel("ncsnpp", cfg, state_dict, lowering_ncsnpp.lower, out_channels=cfg.num_channels)
    else:
        raise NotImplementedError(f'Unknown dataset {config.data.dataset}!')
    return model, img_shape


class RevGuidedDiffusion(torch.nn.Module):
    def __init__(self, args, config, device=None, state_dict=None):
        """Same arguments as the reference (L151); `state_dict` optionally supplies the UNet weights
        (reference parameter names) instead of the pretrained checkpoint files."""
        super().__init__()
        self.args = args
        self.config = config
        if device is None:
            device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)

        model, img_shape = build_score_model(config, state_dict)
        model.eval()
        self.model = model
        self.img_shape = img_shape
        self.rev_vpsde = RevVPSDE(model=model, score_type=args.score_type, img_shape=img_shape, model_kwargs=None)
        self.betas = self.rev_vpsde.discrete_betas.float().to(self.device)
        self._tables = {}
        self.sample_offset = 0      # global index of this shard's first sample (multi-GPU sharding)
        self.last_seed = None

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
        assert isinstance(img, torch.Tensor)
        if torch.is_grad_enabled() and img.requires_grad:
            raise NotImplementedError("diffpure_b200: backward through the purification loop (torchsde adjoint) is "
                                      "not implemented; wrap the call in torch.no_grad() / detach the input")
        batch_size = img.shape[0]
        if tag is None:
            tag = 'rnd' + str(random.randint(0, 10000))
        out_dir = os.path.join(self.args.log_dir, 'bs' + str(bs_id) + '_' + tag)
        assert img.ndim == 4, img.ndim
        dev = self.device if self.device.type == "cuda" else img.device
        img = img.to(dev)
        x0 = img
        save = bs_id < 2 and getattr(self.args, "save_images", True)
        if save:
            import torchvision.utils as tvu
            os.makedirs(out_dir, exist_ok=True)
            tvu.save_image((x0 + 1) * 0.5, os.path.join(out_dir, 'original_input.png'))

        eng = self.model.engine_for(batch_size, dev)
        cond, coef = self._tables_for(self.args.t)
        xs = []
        for it in range(self.args.sample_step):
            e = torch.randn_like(x0) if init_noise is None else init_noise.to(dev)    # L217
            total_noise_levels = self.args.t
            if self.args.rand_t:
                total_noise_levels = self.args.t + np.random.randint(-self.args.t_delta, self.args.t_delta)
                print(f'total_noise_levels: {total_noise_levels}')
            sx, se = schedule.vpsde_forward_scales(total_noise_levels)                  # L222-223
            if save:
                import torchvision.utils as tvu
                tvu.save_image((x0 * sx + e * se + 1) * 0.5, os.path.join(out_dir, f'init_{it}.png'))
            call_seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed) + it
            self.last_seed = call_seed
            x0 = eng.purify(x0, cond, coef, sx, se, init_noise=e, step_noise=step_noise, seed=call_seed,
                            sample_offset=self.sample_offset)                           # L228-239
            if save:
                import torchvision.utils as tvu
                torch.save(x0, os.path.join(out_dir, f'samples_{it}.pth'))
                tvu.save_image((x0 + 1) * 0.5, os.path.join(out_dir, f'samples_{it}.png'))
            xs.append(x0)
        return torch.cat(xs, dim=0)
