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
        # The reference builds a SpacedDiffusion from these fields (script_util.py:82-132, respace.py:63-99); the same chain
        # is rebuilt here as host tables (schedule.GuidedChain) -- configs/imagenet.yml ships the full 1000-step linear
        # chain with rescaled timesteps, and invites changing timestep_respacing.
        m = config.model
        self._chain_kw = dict(n=int(getattr(m, "diffusion_steps", 1000)),
                              noise_schedule=getattr(m, "noise_schedule", "linear"),
                              timestep_respacing=str(getattr(m, "timestep_respacing", "") or ""),
                              rescale_timesteps=bool(getattr(m, "rescale_timesteps", True)))
        chain = schedule.GuidedChain(**self._chain_kw)
        if state_dict is None:
            state_dict = torch.load(f'{model_dir}/256x256_diffusion_uncond.pt', map_location='cpu')
        self.model = ScoreModel("adm", cfg, state_dict, lowering_adm.lower, out_channels=6).eval()
        self.num_timesteps = chain.num_timesteps
        self.diffusion = SimpleNamespace(betas=chain.betas, num_timesteps=chain.num_timesteps,
                                         timestep_map=chain.timestep_map.tolist(),
                                         rescale_timesteps=chain.rescale, original_num_steps=chain.n_base)
        self.betas = torch.from_numpy(chain.betas).float().to(self.device)

    def image_editing_sample(self, img, bs_id=0, tag=None, init_noise=None, step_noise=None, seed=None):
        with torch.no_grad():
            x0, dev, dump = self._open(img, bs_id, tag)
            eng = self.model.engine_for(x0.shape[0], dev)
            cond, coef, sx, se = schedule.guided_tables(self.args.t, **self._chain_kw)

            def one_pass(it, x):
                e = self._init_noise(x, init_noise, dev)
                if dump.on:
                    dump.image(f'init_{it}.png', x * sx + e * se)
                return eng.purify(x, cond, coef, sx, se, update_kind=_lib.DP_UPDATE_LEARNED_RANGE, init_noise=e,
                                  step_noise=step_noise, seed=self._call_seed(seed, it),
                                  sample_offset=self.sample_offset, **self._fuse_kw)

            return self._passes(x0, dump, one_pass)
