"""GPU (-m gpu): ADM and CelebA-HQ DDPM engines through the C ABI vs the golden vectors from the reference.
Tolerances as in tests/test_gpu_parity.py (eval 2e-2, trajectory 5e-3)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import adm as A, ddpm_unet as D, weights

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
TOL_EVAL, TOL_TRAJ = 2e-2, 5e-3


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def load(name):
    return {k: torch.from_numpy(v) if v.ndim else v for k, v in np.load(os.path.join(G, name)).items()}


def test_adm_eval_and_guided_chain_golden():
    from diffpure_b200 import lib, lowering_adm as LA, schedule
    from diffpure_b200.engine import Engine
    cfg = A.tiny_cfg(64, 64, (1, 2, 3, 4), 1, (32, 16, 8))
    d = load("adm_tiny.npz")
    sd = weights.make_state_dict(A.param_shapes(cfg), seed=int(d["seed"]))
    eng = Engine(LA.lower(cfg, sd, 2), device=0)
    y = eng.unet_forward(d["x"].cuda(), d["t"].float().cuda()).cpu()
    assert y.shape == (2, 6, 64, 64)
    assert rel(y, d["y"]) < TOL_EVAL, rel(y, d["y"])
    cond, coef, sx, se = schedule.guided_tables(int(d["t_levels"]))
    out = eng.purify(d["x0"].cuda(), cond, coef, sx, se, update_kind=lib.DP_UPDATE_LEARNED_RANGE,
                     init_noise=d["e0"].cuda(), step_noise=d["z"].cuda()).cpu()
    eng.close()
    assert rel(out, d["loop_out"]) < TOL_TRAJ, rel(out, d["loop_out"])


def test_celeba_eval_and_chain_golden():
    from diffpure_b200 import lowering_ddpm as LD, schedule
    from diffpure_b200.engine import Engine
    cfg = D.tiny_cfg(32, 64, (1, 2, 2), 1, (16,))
    d = load("celeba_tiny.npz")
    sd = weights.make_state_dict(D.param_shapes(cfg), seed=int(d["seed"]))
    eng = Engine(LD.lower(cfg, sd, 2), device=0)
    y = eng.unet_forward(d["x"].cuda(), d["t"].float().cuda()).cpu()
    assert rel(y, d["y"]) < TOL_EVAL, rel(y, d["y"])
    cond, coef, sx, se = schedule.ddpm_tables(int(d["t_levels"]))
    out = eng.purify(d["x0"].cuda(), cond, coef, sx, se, init_noise=d["e0"].cuda(), step_noise=d["z"].cuda()).cpu()
    eng.close()
    assert rel(out, d["loop_out"]) < TOL_TRAJ, rel(out, d["loop_out"])


def test_guided_and_ddpm_runner_api():
    """GuidedDiffusion / Diffusion keep the reference's constructor + image_editing_sample signatures."""
    from diffpure_b200.runners.diffpure_guided import GuidedDiffusion
    from diffpure_b200.runners.diffpure_ddpm import Diffusion
    args = SimpleNamespace(t=3, sample_step=1, log_dir="/tmp/dp_test_logs", save_images=False)
    acfg = A.tiny_cfg(64, 64, (1, 2, 3, 4), 1, (32, 16, 8))
    config = SimpleNamespace(data=SimpleNamespace(dataset="ImageNet"),
                             model=SimpleNamespace(image_size=64, num_channels=64, num_res_blocks=1,
                                                   attention_resolutions="32,16,8", num_head_channels=64,
                                                   use_scale_shift_norm=True, resblock_updown=True, learn_sigma=True,
                                                   class_cond=False, diffusion_steps=1000, channel_mult=""))
    d = load("adm_tiny.npz")
    r = GuidedDiffusion(args, config, device=torch.device("cuda:0"),
                        state_dict=weights.make_state_dict(A.param_shapes(acfg), seed=int(d["seed"])))
    out = r.image_editing_sample(d["x0"].cuda(), bs_id=3, tag="x", init_noise=d["e0"].cuda(), step_noise=d["z"].cuda())
    assert rel(out.cpu(), d["loop_out"]) < TOL_TRAJ
    r.model.release()

    ccfg = D.tiny_cfg(32, 64, (1, 2, 2), 1, (16,))
    config = SimpleNamespace(data=SimpleNamespace(dataset="CelebA_HQ", image_size=32),
                             model=SimpleNamespace(ch=64, out_ch=3, ch_mult=[1, 2, 2], num_res_blocks=1,
                                                   attn_resolutions=[16], in_channels=3, resamp_with_conv=True,
                                                   var_type="fixedsmall"),
                             diffusion=SimpleNamespace(beta_start=1e-4, beta_end=2e-2, num_diffusion_timesteps=1000))
    d = load("celeba_tiny.npz")
    args = SimpleNamespace(t=int(d["t_levels"]), sample_step=1, log_dir="/tmp/dp_test_logs", save_images=False)
    r = Diffusion(args, config, device=torch.device("cuda:0"),
                  state_dict=weights.make_state_dict(D.param_shapes(ccfg), seed=int(d["seed"])))
    out = r.image_editing_sample(d["x0"].cuda(), bs_id=3, tag="x", init_noise=d["e0"].cuda(), step_noise=d["z"].cuda())
    assert rel(out.cpu(), d["loop_out"]) < TOL_TRAJ
    with pytest.raises(ValueError):
        Diffusion(args, SimpleNamespace(data=SimpleNamespace(dataset="LSUN"), model=config.model,
                                        diffusion=config.diffusion), device=torch.device("cuda:0"), state_dict={})
    r.model.release()


@pytest.mark.parametrize("which", ["adm", "celeba"])
def test_full_size_256_eval_vs_oracle(which):
    """BASELINE-size 256x256 UNets (ADM 552.8 M / CelebA-HQ 113.7 M parameters), one evaluation at B=1 vs the CPU oracle.
    At 256x256 every 3x3 conv has >= 512 M tiles, so this is also the end-to-end parity check of the CTA-pair
    (cta_group::2) GEMM tiles inside a whole network."""
    from diffpure_b200 import synthetic
    from diffpure_b200.engine import Engine
    if which == "adm":
        from diffpure_b200 import lowering_adm as L
        cfg, ocfg, O = L.imagenet_cfg(), A.IMAGENET_CFG, A
    else:
        from diffpure_b200 import lowering_ddpm as L
        cfg, ocfg, O = L.celeba_cfg(), D.CELEBA_CFG, D
    sd = synthetic.random_state_dict(L.param_shapes(cfg), seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    t = torch.tensor([77.0])
    with torch.no_grad():
        y = O.forward(ocfg, sd, x, t if which == "adm" else t.long())
    eng = Engine(L.lower(cfg, sd, 1), device=0)
    n_pair = eng.pair_gemms
    yg = eng.unet_forward(x.cuda(), t.cuda()).cpu()
    eng.close()
    assert n_pair > 0
    assert rel(yg, y) < TOL_EVAL, rel(yg, y)
