"""GuidedDiffusion with the reference's API (runners/diffpure_guided.py:17-89), backed by the sm_100a engine.

The ancestral chain `for i in reversed(range(t)): x = diffusion.p_sample(model, x, t, clip_denoised=True)` (L68-75)
runs on the device: ADM UNet evaluation + learned-range variance / x0-clamp / posterior-mean update fused into the
output conv's epilogue (`DP_UPDATE_LEARNED_RANGE`). Runs under no_grad like the reference (L42).
"""
import os
import random
from types import SimpleNamespace

import numpy as np
import torch

from .. import lib as _lib
from .. import lowering_adm, schedule
from ..model import ScoreModel


class GuidedDiffusion(torch.nn.Module):
    def __init__(self, args, config, device=None, model_dir='pretrained/guided_diffusion', state_dict=None):
        super().__init__()
        self.args = args
        self.config = config
        if device is None:
            device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        cfg = lowering_adm.cfg_from_reference(config)
        if state_dict is None:
            state_dict = torch.load(f'{model_dir}/256x256_diffusion_uncond.pt', map_location='cpu')
        self.model = ScoreModel("adm", cfg, state_dict, lowering_adm.lower, out_channels=6).eval()
        self.num_timesteps = int(getattr(config.model, "diffusion_steps", 1000))
        _, _, _, _ = schedule.guided_tables(1, self.num_timesteps)
        base = schedule._linear_betas64(self.num_timesteps)
        self.diffusion = SimpleNamespace(betas=base, num_timesteps=self.num_timesteps)
        self.betas = torch.from_numpy(base).float().to(self.device)
        self.sample_offset = 0
        self.last_seed = None

    def image_editing_sample(self, img, bs_id=0, tag=None, init_noise=None, step_noise=None, seed=None):
        with torch.no_grad():
            assert isinstance(img, torch.Tensor)
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
            cond, coef, sx, se = schedule.guided_tables(self.args.t, self.num_timesteps)
            xs = []
            for it in range(self.args.sample_step):
                e = torch.randn_like(x0) if init_noise is None else init_noise.to(dev)
                if save:
                    import torchvision.utils as tvu
                    tvu.save_image((x0 * sx + e * se + 1) * 0.5, os.path.join(out_dir, f'init_{it}.png'))
                call_seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed) + it
                self.last_seed = call_seed
                x0 = eng.purify(x0, cond, coef, sx, se, update_kind=_lib.DP_UPDATE_LEARNED_RANGE, init_noise=e,
                                step_noise=step_noise, seed=call_seed, sample_offset=self.sample_offset)
                if save:
                    import torchvision.utils as tvu
                    torch.save(x0, os.path.join(out_dir, f'samples_{it}.pth'))
                    tvu.save_image((x0 + 1) * 0.5, os.path.join(out_dir, f'samples_{it}.png'))
                xs.append(x0)
            return torch.cat(xs, dim=0)
