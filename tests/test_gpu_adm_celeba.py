"""GPU (-m gpu): ADM and CelebA-HQ DDPM engines through the C ABI vs the golden vectors from the reference.
Tolerances as in tests/test_gpu_parity.py (eval 2e-2, trajectory 5e-3)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from golden_inputs import ADM_TINY_REF_CONFIG, adm_vpsde_inputs, fullsize_chain_inputs, fullsize_eval_inputs, \
    respaced_chain_inputs
from oracle import adm as A, ddpm_unet as D, sde as OS, weights

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


def test_guided_runner_respaced_chain_golden():
    """config.model.timestep_respacing = '100,50,25': GuidedDiffusion on the engine vs the chain the reference's own
    SpacedDiffusion.p_sample produced (tests/golden/guided_schedules.npz; oracle/make_golden.py:golden_guided_schedules)."""
    from diffpure_b200.runners.diffpure_guided import GuidedDiffusion
    d = np.load(os.path.join(G, "guided_schedules.npz"))
    mc = dict(ADM_TINY_REF_CONFIG, timestep_respacing=str(d["chain_respacing"]))
    args = SimpleNamespace(t=int(d["chain_t_levels"]), sample_step=1, log_dir="/tmp/dp_test_logs", save_images=False)
    config = SimpleNamespace(data=SimpleNamespace(dataset="ImageNet"), model=SimpleNamespace(**mc))
    acfg = A.tiny_cfg(64, 64, (1, 2, 3, 4), 1, (32, 16, 8))
    r = GuidedDiffusion(args, config, device=torch.device("cuda:0"),
                        state_dict=weights.make_state_dict(A.param_shapes(acfg), seed=int(d["chain_seed"])))
    x0, e0 = respaced_chain_inputs(int(d["chain_input_seed"]))
    out = r.image_editing_sample(x0.cuda(), bs_id=3, tag="rs", init_noise=e0.cuda(),
                                 step_noise=torch.from_numpy(d["chain_z"]).cuda())
    r.model.release()
    want = torch.from_numpy(d["chain_out"])
    assert rel(out.cpu(), want) < TOL_TRAJ, rel(out.cpu(), want)


@pytest.mark.parametrize("which", ["adm", "celeba"])
def test_full_size_256_eval_vs_oracle(which):
    """BASELINE-size 256x256 UNets (ADM 552.8 M / CelebA-HQ 113.7 M parameters), one evaluation at B=1 vs the CPU oracle
    (precomputed: tests/golden/fullsize_oracle.npz). At 256x256 every 3x3 conv has >= 512 M tiles, so this is also the
    end-to-end parity check of the CTA-pair (cta_group::2) GEMM tiles inside a whole network."""
    import fullsize as F
    x, t = fullsize_eval_inputs()
    y = F.oracle_results()[f"{which}_eval"]
    eng = F.engine(which, 1)
    n_pair = eng.pair_gemms
    yg = eng.unet_forward(x.cuda(), t.cuda()).cpu()
    eng.close()
    assert n_pair > 0
    assert torch.isfinite(yg).all()
    assert rel(F.sparse(yg), y) < TOL_EVAL, rel(F.sparse(yg), y)


def test_adm_on_the_vpsde_path_runner_golden():
    """The canonical ImageNet configuration (run_scripts/imagenet/run_in_rand_inf.sh:12-24: --diffusion_type sde, default
    score_type 'guided_diffusion'): RevGuidedDiffusion(dataset='ImageNet') on the engine vs the fixture produced by the
    reference's RevVPSDE around its own ADM UNet (runners/diffpure_sde.py:101-112,160-170)."""
    from diffpure_b200.runners.diffpure_sde import RevGuidedDiffusion
    d = load("adm_tiny_vpsde.npz")
    t_star = int(d["t_star"])
    x0, e0, z = adm_vpsde_inputs(d["input_seed"], t_star)
    acfg = A.tiny_cfg(64, 64, (1, 2, 3, 4), 1, (32, 16, 8))
    sd = weights.make_state_dict(A.param_shapes(acfg), seed=int(d["seed"]))
    args = SimpleNamespace(t=t_star, rand_t=False, t_delta=15, use_bm=False, score_type="guided_diffusion",
                           sample_step=1, log_dir="/tmp/dp_test_logs", save_images=False)
    config = SimpleNamespace(data=SimpleNamespace(dataset="ImageNet"), model=SimpleNamespace(**ADM_TINY_REF_CONFIG))
    r = RevGuidedDiffusion(args, config, device=torch.device("cuda:0"), state_dict=sd)
    with torch.no_grad():
        out = r.image_editing_sample(x0.cuda(), bs_id=3, tag="x", init_noise=e0.cuda(), step_noise=z.cuda()).cpu()
    assert out.shape == (2, 3, 64, 64)
    assert rel(out, d["loop_out"]) < TOL_TRAJ, rel(out, d["loop_out"])
    # the SDE-object protocol on the same engine: f at the first grid point vs the reference's
    xs = OS.forward_diffuse(x0, e0, t_star)
    f0 = r.rev_vpsde.f(OS.time_grid(t_star)[0].cuda(), xs.cuda().reshape(2, -1)).reshape(2, 3, 64, 64).cpu()
    assert rel(f0, d["f0"]) < TOL_EVAL, rel(f0, d["f0"])
    r.model.release()


@pytest.mark.parametrize("which", ["adm_vpsde", "adm_guided", "celeba"])
def test_full_size_256_three_step_chain_vs_oracle(which):
    """BASELINE-size 256x256 models, 3 steps of the real schedule (the first three of the 150- / 100-step chains) at B=1
    with injected noise vs the CPU oracle loop (precomputed: tests/golden/fullsize_oracle.npz)."""
    import fullsize as F
    from diffpure_b200 import lib, schedule
    x0, e0, z = fullsize_chain_inputs()
    kind = lib.DP_UPDATE_LINEAR
    if which == "adm_vpsde":
        cond, coef = schedule.vpsde_tables(150, "guided_diffusion")
        sx, se = schedule.vpsde_forward_scales(150)
    elif which == "adm_guided":
        cond, coef, sx, se = schedule.guided_tables(150)
        kind = lib.DP_UPDATE_LEARNED_RANGE
    else:
        cond, coef, sx, se = schedule.ddpm_tables(100)
    eng = F.engine("celeba" if which == "celeba" else "adm", 1)
    out = eng.purify(x0.cuda(), cond[:3], coef[:3], sx, se, update_kind=kind, init_noise=e0.cuda(),
                     step_noise=z.cuda()).cpu()
    eng.close()
    want = F.oracle_results()[f"chain_{which}"]
    assert torch.isfinite(out).all()
    assert rel(F.sparse(out), want) < TOL_TRAJ, rel(F.sparse(out), want)


@pytest.mark.parametrize("which,B", [("adm", 32), ("celeba", 16)])
def test_full_size_256_benchmark_batch_matches_b1_engine(which, B):
    """At the benchmarked batch (ADM 32, CelebA 16) the tile choices differ from B=1 (CTA pairs everywhere, BN=256):
    3 steps of the chain at batch B vs the same samples through B=1 engines -- same arithmetic, different tiling."""
    import fullsize as F
    from diffpure_b200 import schedule
    if which == "adm":
        cond, coef = schedule.vpsde_tables(150, "guided_diffusion")
        sx, se = schedule.vpsde_forward_scales(150)
    else:
        cond, coef, sx, se = schedule.ddpm_tables(100)
    g = torch.Generator().manual_seed(22)
    x0 = torch.rand(B, 3, 256, 256, generator=g) * 2 - 1
    engB = F.engine(which, B)
    big = engB.purify(x0.cuda(), cond[:3], coef[:3], sx, se, seed=5, sample_offset=0).cpu()
    engB.close()
    eng1 = F.engine(which, 1)
    for i in (0, B - 1):
        one = eng1.purify(x0[i:i + 1].cuda(), cond[:3], coef[:3], sx, se, seed=5, sample_offset=i).cpu()
        assert rel(big[i:i + 1], one) < 1e-3, (i, rel(big[i:i + 1], one))
    eng1.close()
    assert torch.isfinite(big).all()
