"""Diffusion (CelebA-HQ DDPM) with the reference's API (runners/diffpure_ddpm.py:57-142), backed by the sm_100a engine.

`image_editing_denoising_step_flexible_mask` (L37-54) is linear in (x, eps, z), so the whole chain runs on the device
with the fused `DP_UPDATE_LINEAR` update behind the output conv.
"""
import numpy as np
import torch

from .. import lowering_ddpm, schedule
from ..model import ScoreModel
from ._common import PurifyRunner


def get_beta_schedule(*, beta_start, beta_end, num_diffusion_timesteps):
    betas = np.linspace(beta_start, beta_end, num_diffusion_timesteps, dtype=np.float64)
    assert betas.shape == (num_diffusion_timesteps,)
    return betas


class Diffusion(PurifyRunner):
    device_from_input = True

    def __init__(self, args, config, device=None, state_dict=None):
        super().__init__()
        self._setup(args, config, device)
        print("Loading model")
        if self.config.data.dataset != "CelebA_HQ":
            raise ValueError
        if state_dict is None:
            url = "https://image-editing-test-12345.s3-us-west-2.amazonaws.com/checkpoints/celeba_hq.ckpt"
            state_dict = torch.hub.load_state_dict_from_url(url, map_location='cpu')
        cfg = lowering_ddpm.cfg_from_reference(config)
        self.model = ScoreModel("ddpm", cfg, state_dict, lowering_ddpm.lower, out_channels=cfg.out_ch).eval()
        self.model_var_type = config.model.var_type
        d = config.diffusion
        self._sched = (d.beta_start, d.beta_end, d.num_diffusion_timesteps)
        betas = get_beta_schedule(beta_start=d.beta_start, beta_end=d.beta_end,
                                  num_diffusion_timesteps=d.num_diffusion_timesteps)
        self.betas = torch.from_numpy(betas).float()
        self.num_timesteps = betas.shape[0]

    def image_editing_sample(self, img=None, bs_id=0, tag=None, init_noise=None, step_noise=None, seed=None):
        with torch.no_grad():
            x0, dev, dump = self._open(img, bs_id, tag)
            eng = self.model.engine_for(x0.shape[0], dev)
            cond, coef, sx, se = schedule.ddpm_tables(self.args.t, *self._sched, var_type=self.model_var_type)

            def one_pass(it, x):
                e = self._init_noise(x, init_noise, dev)
                if dump.on:
                    dump.image(f'init_{it}.png', x * sx + e * se)
                return eng.purify(x, cond, coef, sx, se, init_noise=e, step_noise=step_noise,
                                  seed=self._call_seed(seed, it), sample_offset=self.sample_offset, **self._fuse_kw)

            return self._passes(x0, dump, one_pass)
