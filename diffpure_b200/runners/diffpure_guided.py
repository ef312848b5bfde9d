"""GuidedDiffusion with the reference's API (runners/diffpure_guided.py:17-89), backed by the sm_100a engine.

The ancestral chain `for i in reversed(range(t)): x = diffusion.p_sample(model, x, t, clip_denoised=True)` (L68-75)
runs on the device: ADM UNet evaluation + learned-range variance / x0-clamp / posterior-mean update fused behind the
output conv (`DP_UPDATE_LEARNED_RANGE`). Runs under no_grad like the reference (L42).
"""
from types import SimpleNamespace

import torch

from .. import lib as _lib
from .. import lowering_adm, schedule
from ..model import ScoreModel
from ._common import PurifyRunner


class GuidedDiffusion(PurifyRunner):
    def __init__(self, args, config, device=None, model_dir='pretrained/guided_diffusion', state_dict=None):
        super().__init__()
        self._setup(args, config, device)
        cfg = lowering_adm.cfg_from_reference(config)
        # The reference builds a SpacedDiffusion from these fields (script_util.py:82-132, respace.py:71-99); this runner
        # implements the configuration DiffPure ships (configs/imagenet.yml: full 1000-step linear chain, rescaled
        # timesteps) and refuses anything else instead of silently running a different chain.
        m = config.model
        n_steps = int(getattr(m, "diffusion_steps", 1000))
        if str(getattr(m, "timestep_respacing", "") or "") not in ("", str(n_steps)):
            raise NotImplementedError(f"timestep_respacing={m.timestep_respacing!r} is not supported (only the full chain)")
        if getattr(m, "noise_schedule", "linear") != "linear":
            raise NotImplementedError(f"noise_schedule={m.noise_schedule!r} is not supported (only 'linear')")
        if not getattr(m, "rescale_timesteps", True):
            raise NotImplementedError("rescale_timesteps=False is not supported")
        if state_dict is None:
            state_dict = torch.load(f'{model_dir}/256x256_diffusion_uncond.pt', map_location='cpu')
        self.model = ScoreModel("adm", cfg, state_dict, lowering_adm.lower, out_channels=6).eval()
        self.num_timesteps = int(getattr(config.model, "diffusion_steps", 1000))
        base = schedule._linear_betas64(self.num_timesteps)
        self.diffusion = SimpleNamespace(betas=base, num_timesteps=self.num_timesteps)
        self.betas = torch.from_numpy(base).float().to(self.device)

    def image_editing_sample(self, img, bs_id=0, tag=None, init_noise=None, step_noise=None, seed=None):
        with torch.no_grad():
            x0, dev, dump = self._open(img, bs_id, tag)
            eng = self.model.engine_for(x0.shape[0], dev)
            cond, coef, sx, se = schedule.guided_tables(self.args.t, self.num_timesteps)

            def one_pass(it, x):
                e = self._init_noise(x, init_noise, dev)
                if dump.on:
                    dump.image(f'init_{it}.png', x * sx + e * se)
                return eng.purify(x, cond, coef, sx, se, update_kind=_lib.DP_UPDATE_LEARNED_RANGE, init_noise=e,
                                  step_noise=step_noise, seed=self._call_seed(seed, it),
                                  sample_offset=self.sample_offset, **self._fuse_kw)

            return self._passes(x0, dump, one_pass)
