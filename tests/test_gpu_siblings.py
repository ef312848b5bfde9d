"""GPU (-m gpu): the ODE and Langevin-SDE sibling runners (same engine, different per-step update) vs the oracle loops."""
from types import SimpleNamespace

import pytest
import torch

from oracle import ncsnpp as O, sde as OS, weights

pytestmark = pytest.mark.gpu
TOL_TRAJ = 5e-3


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def _config():
    return SimpleNamespace(data=SimpleNamespace(dataset="CIFAR10", image_size=32, num_channels=3),
                           model=SimpleNamespace(name="ncsnpp", resblock_type="biggan", fir=False, skip_rescale=True,
                                                 progressive="none", progressive_input="none",
                                                 embedding_type="positional", conditional=True, nonlinearity="swish",
                                                 nf=64, ch_mult=[1, 2, 2], num_res_blocks=1, attn_resolutions=[16]))


def test_ode_runner_vs_oracle():
    from diffpure_b200.runners.diffpure_ode import OdeGuidedDiffusion, VPODE
    cfg = O.tiny_cfg(64, (1, 2, 2), 1, (16,), 32)
    sd = weights.make_state_dict(O.param_shapes(cfg), seed=1)
    unet = lambda x, t: O.forward(cfg, sd, x, t)  # noqa: E731
    g = torch.Generator().manual_seed(21)
    img = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    e0 = torch.randn(2, 3, 32, 32, generator=g)
    args = SimpleNamespace(t=20, step_size=1e-3, score_type="score_sde", sample_step=1, log_dir="/tmp/dp_test_logs",
                           save_images=False)
    r = OdeGuidedDiffusion(args, _config(), device=torch.device("cuda:0"), state_dict=sd)
    assert isinstance(r.vpode, VPODE)
    out = r.image_editing_sample(img.cuda(), bs_id=9, tag="o", init_noise=e0.cuda()).cpu()
    ref = OS.purify_ode(unet, img, 20, e0, step_size=1e-3)
    assert rel(out, ref) < TOL_TRAJ, rel(out, ref)
    # torchdiffeq protocol of VPODE.forward vs the oracle integrand
    dx = r.vpode(torch.tensor(0.07).cuda(), (img.cuda().reshape(2, -1),))[0].reshape(2, 3, 32, 32).cpu()
    assert rel(dx, OS.vpode_f(unet, "score_sde", torch.tensor(0.07), img)) < 2e-2
    r.model.release()


def test_ldsde_runner_vs_oracle():
    from diffpure_b200.runners.diffpure_ldsde import LDGuidedDiffusion
    cfg = O.tiny_cfg(64, (1, 2, 2), 1, (16,), 32)
    sd = weights.make_state_dict(O.param_shapes(cfg), seed=1)
    unet = lambda x, t: O.forward(cfg, sd, x, t)  # noqa: E731
    g = torch.Generator().manual_seed(22)
    img = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    n = OS.num_steps_ldsde(100)
    z = torch.randn(n, 2, 3, 32, 32, generator=g)
    args = SimpleNamespace(t=100, sigma2=1e-3, lambda_ld=1e-2, eta=5, score_type="score_sde", sample_step=1,
                           log_dir="/tmp/dp_test_logs", save_images=False, use_bm=False)
    r = LDGuidedDiffusion(args, _config(), device=torch.device("cuda:0"), state_dict=sd)
    out = r.image_editing_sample(img.cuda(), bs_id=9, tag="l", step_noise=z.cuda()).cpu()
    ref = OS.purify_ldsde(unet, img, 100, z, sigma2=1e-3, lambda_ld=1e-2, eta=5.0)
    assert rel(out, ref) < TOL_TRAJ, rel(out, ref)
    # generated noise path: deterministic in the seed
    a = r.image_editing_sample(img.cuda(), bs_id=9, tag="l", seed=5)
    b = r.image_editing_sample(img.cuda(), bs_id=9, tag="l", seed=5)
    assert torch.equal(a, b)
    r.model.release()
