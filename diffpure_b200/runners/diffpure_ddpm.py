"""Diffusion (CelebA-HQ DDPM) with the reference's API (runners/diffpure_ddpm.py:57-142), backed by the sm_100a engine.

`image_editing_denoising_step_flexible_mask` (L37-54) is linear in (x, eps, z), so the whole chain runs on the device
with the fused `DP_UPDATE_LINEAR` epilogue of the output conv.
"""
import os
import random

import numpy as np
import torch

from .. import lowering_ddpm, schedule
from ..model import ScoreModel


def get_beta_schedule(*, beta_start, beta_end, num_diffusion_timesteps):
    betas = np.linspace(beta_start, beta_end, num_diffusion_timesteps, dtype=np.float64)
    assert betas.shape == (num_diffusion_timesteps,)
    return betas


class Diffusion(torch.nn.Module):
    def __init__(self, args, config, device=None, state_dict=None):
        super().__init__()
        self.args = args
        self.config = config
        if device is None:
            device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        print("Loading model")
        if self.config.data.dataset == "CelebA_HQ":
            url = "https://image-editing-test-12345.s3-us-west-2.amazonaws.com/checkpoints/celeba_hq.ckpt"
        else:
            raise ValueError
        if state_dict is None:
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
        self.sample_offset = 0
        self.last_seed = None

    def image_editing_sample(self, img=None, bs_id=0, tag=None, init_noise=None, step_noise=None, seed=None):
        assert isinstance(img, torch.Tensor)
        batch_size = img.shape[0]
        with torch.no_grad():
            if tag is None:
                tag = 'rnd' + str(random.randint(0, 10000))
            out_dir = os.path.join(self.args.log_dir, 'bs' + str(bs_id) + '_' + tag)
            assert img.ndim == 4, img.ndim
            if img.device.type != "cuda":           # the reference takes the device from the input image (L126,129)
                img = img.to(self.device)
            x0 = img
            save = bs_id < 2 and getattr(self.args, "save_images", True)
            if save:
                import torchvision.utils as tvu
                os.makedirs(out_dir, exist_ok=True)
                tvu.save_image((x0 + 1) * 0.5, os.path.join(out_dir, 'original_input.png'))
            eng = self.model.engine_for(batch_size, x0.device)
            cond, coef, sx, se = schedule.ddpm_tables(self.args.t, *self._sched, var_type=self.model_var_type)
            xs = []
            for it in range(self.args.sample_step):
                e = torch.randn_like(x0) if init_noise is None else init_noise.to(x0.device)
                if save:
                    import torchvision.utils as tvu
                    tvu.save_image((x0 * sx + e * se + 1) * 0.5, os.path.join(out_dir, f'init_{it}.png'))
                call_seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed) + it
                self.last_seed = call_seed
                x0 = eng.purify(x0, cond, coef, sx, se, init_noise=e, step_noise=step_noise, seed=call_seed,
                                sample_offset=self.sample_offset)
                if save:
                    import torchvision.utils as tvu
                    torch.save(x0, os.path.join(out_dir, f'samples_{it}.pth'))
                    tvu.save_image((x0 + 1) * 0.5, os.path.join(out_dir, f'samples_{it}.png'))
                xs.append(x0)
            return torch.cat(xs, dim=0)
