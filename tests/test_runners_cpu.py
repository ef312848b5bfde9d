"""CPU: host logic of the five runners (reference API) with the engine replaced by a pure-torch stand-in.

The stand-in implements the documented contract of `Engine.purify` (forward diffusion, then per step one model call at
`cond[k]` and the update selected by `update_kind` with the coefficient row `coef[k]`); the runners' schedules, noise
plumbing, `sample_step` loop and SDE-object protocol classes are then held to the oracle loops with the same stand-in
model. (The CUDA engine itself is held to the same oracle loops by the -m gpu tests.)"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from diffpure_b200 import lib as L
from oracle import adm as A, ddpm_loops as OD, ddpm_unet as D, ncsnpp as O, sde as OS, weights


class StubNet(torch.nn.Module):
    """A small deterministic 'score network': x, t -> Cout channels (stands in for the UNet on both sides)."""

    def __init__(self, cout, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.w = torch.randn(cout, 3, 3, 3, generator=g) * 0.2
        self.cout = cout

    def forward(self, x, t):
        t = torch.as_tensor(t, dtype=torch.float32).reshape(-1, 1, 1, 1)
        y = torch.nn.functional.conv2d(x.float(), self.w, padding=1)
        return torch.tanh(y) * (1.0 + 1e-3 * t)

    # Engine factory interface of diffpure_b200.model.ScoreModel
    vjp_ok = True

    def engine_for(self, batch, device, vjp=False):
        return FakeEngine(self)

    def release(self):
        pass


class FakeEngine:
    def __init__(self, net):
        self.net = net
        self.calls = []

    def purify(self, x0, cond, coef, init_scale_x, init_scale_e, *, update_kind=L.DP_UPDATE_LINEAR, init_noise=None,
               step_noise=None, seed=0, sample_offset=0, anchor=None, states=None):
        self.calls.append(dict(steps=len(cond), kind=update_kind, seed=seed, sample_offset=sample_offset))
        g = torch.Generator().manual_seed(int(seed))
        e = init_noise if init_noise is not None else torch.randn(x0.shape, generator=g)
        x = init_scale_x * x0 + init_scale_e * e
        xi = anchor if anchor is not None else x.clone()
        B = x.shape[0]
        if states is not None:
            states[0] = x
        coef = torch.from_numpy(np.asarray(coef, dtype=np.float32))
        for k in range(len(cond)):
            out = self.net(x, torch.full((B,), float(cond[k])))
            z = step_noise[k] if step_noise is not None else torch.randn(x.shape, generator=g)
            c = coef[k]
            if update_kind == L.DP_UPDATE_LINEAR:
                x = c[0] * x + c[1] * out[:, :3] + c[2] * z
            elif update_kind == L.DP_UPDATE_LINEAR_ANCHORED:
                x = c[0] * x + c[1] * out[:, :3] + c[2] * z + c[3] * xi
            else:                                                  # DP_UPDATE_LEARNED_RANGE, schedule.guided_tables row
                eps, var = out[:, :3], out[:, 3:]
                frac = (var + 1) / 2
                logvar = frac * c[4] + (1 - frac) * c[5]
                x0h = (c[0] * x - c[1] * eps).clamp(-1, 1)
                x = c[2] * x0h + c[3] * x + c[6] * torch.exp(0.5 * logvar) * z
            if states is not None:
                states[k + 1] = x
        return x

    def unet_vjp(self, x, cond, g_out):
        """Engine.unet_vjp contract: J(x, cond)^T g of the first three output channels' network."""
        with torch.enable_grad():
            xx = x.detach().clone().requires_grad_(True)
            y = self.net(xx, cond)[:, :3]
            (gx,) = torch.autograd.grad(y, xx, g_out)
        return gx


def _cifar_config():
    return SimpleNamespace(data=SimpleNamespace(dataset="CIFAR10", image_size=16, num_channels=3),
                           model=SimpleNamespace(name="ncsnpp", resblock_type="biggan", fir=False, skip_rescale=True,
                                                 progressive="none", progressive_input="none",
                                                 embedding_type="positional", conditional=True, nonlinearity="swish",
                                                 nf=64, ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[8]))


def _cifar_sd():
    return weights.make_state_dict(O.param_shapes(O.tiny_cfg(64, (1, 2), 1, (8,), 16)), seed=1)


def _data(B=2, S=16, steps=0, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    e = torch.randn(B, 3, S, S, generator=g)
    z = torch.randn(steps, B, 3, S, S, generator=g) if steps else None
    return x, e, z


def test_revguided_diffusion_flow_and_sde_object():
    from diffpure_b200.runners.diffpure_sde import RevGuidedDiffusion
    args = SimpleNamespace(t=7, rand_t=False, t_delta=3, use_bm=False, score_type="score_sde", sample_step=2,
                           log_dir="/tmp/dp_cpu_logs", save_images=False)
    r = RevGuidedDiffusion(args, _cifar_config(), device="cpu", state_dict=_cifar_sd())
    net = StubNet(3)
    r.model = net
    r.rev_vpsde.model = net
    n = OS.num_steps(7)
    x, e, z = _data(steps=n)
    out = r.image_editing_sample(x, bs_id=3, tag="t", init_noise=e, step_noise=z, seed=5)
    p1 = OS.purify_sde(net, x, 7, e, z)
    p2 = OS.purify_sde(net, p1, 7, e, z)                          # sample_step = 2: the second pass starts from the first
    assert out.shape == (4, 3, 16, 16)
    assert torch.allclose(out[:2], p1, atol=2e-5) and torch.allclose(out[2:], p2, atol=5e-5)
    # torchsde SDE-object protocol
    t = torch.tensor(0.93)
    f = r.rev_vpsde.f(t, x.reshape(2, -1)).reshape(x.shape)
    assert torch.allclose(f, OS.rev_vpsde_f(net, "score_sde", t, x), atol=1e-5)
    assert torch.allclose(r.rev_vpsde.g(t, x.reshape(2, -1))[:, 0], OS.rev_vpsde_g(t, 2), atol=1e-6)
    assert r.rev_vpsde.noise_type == "diagonal" and r.rev_vpsde.sde_type == "ito"
    # white-box gradient: the runner's discrete adjoint == autograd through the oracle loop (eval_sde_adv.py:126-128 use)
    args.sample_step = 1
    xg = x.clone().requires_grad_(True)
    out = r.image_editing_sample(xg, bs_id=3, tag="t", init_noise=e, step_noise=z, seed=5)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(1))
    (gx,) = torch.autograd.grad((out * w).sum(), xg)
    xr = x.clone().requires_grad_(True)
    (gr,) = torch.autograd.grad((OS.purify_sde(net, xr, 7, e, z) * w).sum(), xr)
    assert torch.allclose(gx, gr, atol=1e-5, rtol=1e-4) and gx.abs().max() > 0
    # a network without an input-gradient program must fail loudly, never detach silently
    del StubNet.vjp_ok
    try:
        with pytest.raises(NotImplementedError):
            r.image_editing_sample(x.clone().requires_grad_(True))
    finally:
        StubNet.vjp_ok = True
    # rand_t: the forward-diffusion level is jittered, the reverse grid is not (reference L219-231)
    args.rand_t, args.sample_step = True, 1
    np.random.seed(3)
    lvl = 7 + np.random.randint(-3, 3)
    np.random.seed(3)
    out = r.image_editing_sample(x, bs_id=3, tag="t", init_noise=e, step_noise=z, seed=5)
    assert torch.allclose(out, OS.purify_sde(net, x, 7, e, z, t_level=lvl), atol=2e-5)


def test_revguided_guided_score_type_tables():
    """score_type='guided_diffusion' on the VP-SDE path (ImageNet): integer timesteps, eps of a 6-channel model."""
    from diffpure_b200 import schedule
    net = StubNet(6)
    n = OS.num_steps(5)
    x, e, z = _data(steps=n, seed=4)
    cond, coef = schedule.vpsde_tables(5, "guided_diffusion")
    sx, se = schedule.vpsde_forward_scales(5)
    out = FakeEngine(net).purify(x, cond, coef, sx, se, init_noise=e, step_noise=z)
    assert torch.allclose(out, OS.purify_sde(net, x, 5, e, z, score_type="guided_diffusion"), atol=2e-5)


def test_ode_and_ldsde_runners_flow():
    from diffpure_b200.runners.diffpure_ode import OdeGuidedDiffusion
    from diffpure_b200.runners.diffpure_ldsde import LDGuidedDiffusion
    net = StubNet(3, seed=2)
    args = SimpleNamespace(t=12, step_size=1e-3, score_type="score_sde", sample_step=1, log_dir="/tmp/dp_cpu_logs",
                           save_images=False)
    r = OdeGuidedDiffusion(args, _cifar_config(), device="cpu", state_dict=_cifar_sd())
    r.model = net
    r.vpode.model = net
    x, e, _ = _data(seed=6)
    out = r.image_editing_sample(x, bs_id=4, tag="o", init_noise=e)
    assert torch.allclose(out, OS.purify_ode(net, x, 12, e, step_size=1e-3), atol=2e-5)
    dx = r.vpode(torch.tensor(0.05), (x.reshape(2, -1),))[0].reshape(x.shape)
    assert torch.allclose(dx, OS.vpode_f(net, "score_sde", torch.tensor(0.05), x), atol=1e-5)
    w = torch.randn(x.shape, generator=torch.Generator().manual_seed(2))
    xg, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)        # odeint_adjoint use (L230-238)
    (gx,) = torch.autograd.grad((r.image_editing_sample(xg, bs_id=4, tag="o", init_noise=e) * w).sum(), xg)
    (gr,) = torch.autograd.grad((OS.purify_ode(net, xr, 12, e, step_size=1e-3) * w).sum(), xr)
    assert torch.allclose(gx, gr, atol=1e-5, rtol=1e-4) and gx.abs().max() > 0

    args = SimpleNamespace(t=30, sigma2=1e-3, lambda_ld=1e-2, eta=5, score_type="score_sde", sample_step=2,
                           log_dir="/tmp/dp_cpu_logs", save_images=False, use_bm=False)
    r = LDGuidedDiffusion(args, _cifar_config(), device="cpu", state_dict=_cifar_sd())
    r.model = net
    n = OS.num_steps_ldsde(30)
    z = torch.randn(n, 2, 3, 16, 16, generator=torch.Generator().manual_seed(8))
    out = r.image_editing_sample(x, bs_id=4, tag="l", step_noise=z)
    p1 = OS.purify_ldsde(net, x, 30, z)
    assert torch.allclose(out[:2], p1, atol=5e-5)
    # second pass: starts from the first pass's output but stays anchored to the ORIGINAL image (reference L216)
    xx, g = p1.clone(), float(np.sqrt(1e-2) * 5)
    ts = torch.linspace(1 - 30 / 1000, 1 - 1e-5, 2)
    t, k = ts[0], 0
    while t < ts[-1]:
        tn = torch.minimum(t + 1e-2, ts[-1])
        xx = xx + OS.ldsde_f(net, "score_sde", xx, x, 1e-3, 1e-2) * (tn - t) + g * z[k] * torch.sqrt(tn - t)
        t, k = tn, k + 1
    assert torch.allclose(out[2:], xx, atol=1e-4)
    f = r.ldsde.f(torch.tensor(0.97), p1.reshape(2, -1)).reshape(x.shape)
    assert torch.allclose(f, OS.ldsde_f(net, "score_sde", p1, x, 1e-3, 1e-2), atol=1e-4)
    # gradient through the Langevin loop: the input is both the start state and the anchor of every step
    args.sample_step = 1
    xg, xr = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad((r.image_editing_sample(xg, bs_id=4, tag="l", step_noise=z) * w).sum(), xg)
    (gr,) = torch.autograd.grad((OS.purify_ldsde(net, xr, 30, z) * w).sum(), xr)
    assert torch.allclose(gx, gr, atol=1e-4, rtol=1e-3) and gx.abs().max() > 0


def test_guided_and_celeba_runners_flow():
    from diffpure_b200.runners.diffpure_guided import GuidedDiffusion
    from diffpure_b200.runners.diffpure_ddpm import Diffusion
    args = SimpleNamespace(t=4, sample_step=1, log_dir="/tmp/dp_cpu_logs", save_images=False)
    acfg = A.tiny_cfg(64, 64, (1, 2, 3, 4), 1, (32, 16, 8))
    config = SimpleNamespace(data=SimpleNamespace(dataset="ImageNet"),
                             model=SimpleNamespace(image_size=64, num_channels=64, num_res_blocks=1,
                                                   attention_resolutions="32,16,8", num_head_channels=64,
                                                   use_scale_shift_norm=True, resblock_updown=True, learn_sigma=True,
                                                   class_cond=False, diffusion_steps=1000, channel_mult=""))
    r = GuidedDiffusion(args, config, device="cpu", state_dict=weights.make_state_dict(A.param_shapes(acfg), seed=0))
    net6 = StubNet(6, seed=3)
    r.model = net6
    x, e, z = _data(steps=4, seed=9)
    out = r.image_editing_sample(x, bs_id=5, tag="g", init_noise=e, step_noise=z)
    assert torch.allclose(out, OD.purify_guided(net6, x, 4, e, z), atol=2e-5)
    # timestep_respacing / noise_schedule / rescale_timesteps build the reference's SpacedDiffusion chain (respace.py:63-136)
    config.model.timestep_respacing, config.model.noise_schedule, config.model.rescale_timesteps = "100,50,25", "cosine", False
    r = GuidedDiffusion(args, config, device="cpu", state_dict=weights.make_state_dict(A.param_shapes(acfg), seed=0))
    r.model = net6
    assert r.num_timesteps == 175 and r.diffusion.timestep_map[:4] == [0, 3, 7, 10] and len(r.betas) == 175
    out = r.image_editing_sample(x, bs_id=5, tag="g", init_noise=e, step_noise=z)
    ref = OD.purify_guided(net6, x, 4, e, z, noise_schedule="cosine", timestep_respacing="100,50,25", rescale_timesteps=False)
    assert torch.allclose(out, ref, atol=2e-5) and not torch.allclose(ref, OD.purify_guided(net6, x, 4, e, z), atol=1e-3)

    ccfg = D.tiny_cfg(32, 64, (1, 2, 2), 1, (16,))
    config = SimpleNamespace(data=SimpleNamespace(dataset="CelebA_HQ", image_size=32),
                             model=SimpleNamespace(ch=64, out_ch=3, ch_mult=[1, 2, 2], num_res_blocks=1,
                                                   attn_resolutions=[16], in_channels=3, resamp_with_conv=True,
                                                   var_type="fixedsmall"),
                             diffusion=SimpleNamespace(beta_start=1e-4, beta_end=2e-2, num_diffusion_timesteps=1000))
    r = Diffusion(args, config, device="cpu", state_dict=weights.make_state_dict(D.param_shapes(ccfg), seed=0))
    net3 = StubNet(3, seed=4)
    r.model = net3
    out = r.image_editing_sample(x, bs_id=5, tag="c", init_noise=e, step_noise=z)
    assert torch.allclose(out, OD.purify_celeba(net3, x, 4, e, z), atol=2e-5)
    with pytest.raises(ValueError):
        Diffusion(args, SimpleNamespace(data=SimpleNamespace(dataset="LSUN"), model=config.model,
                                        diffusion=config.diffusion), device="cpu", state_dict={})


def test_image_dumps_follow_the_reference_layout(tmp_path):
    """bs_id < 2 writes original_input.png / init_*.png / samples_*.{pth,png} under log_dir/bs{bs_id}_{tag} (L210-243)."""
    import os
    from diffpure_b200.runners.diffpure_sde import RevGuidedDiffusion
    args = SimpleNamespace(t=2, rand_t=False, t_delta=1, use_bm=False, score_type="score_sde", sample_step=1,
                           log_dir=str(tmp_path))
    r = RevGuidedDiffusion(args, _cifar_config(), device="cpu", state_dict=_cifar_sd())
    net = StubNet(3)
    r.model = net
    r.rev_vpsde.model = net
    x, e, _ = _data()
    r.image_editing_sample(x, bs_id=1, tag="dump", init_noise=e, seed=1)
    d = os.path.join(str(tmp_path), "bs1_dump")
    assert sorted(os.listdir(d)) == ["init_0.png", "original_input.png", "samples_0.png", "samples_0.pth"]
    r.image_editing_sample(x, bs_id=2, tag="nodump", init_noise=e, seed=1)
    assert not os.path.exists(os.path.join(str(tmp_path), "bs2_nodump"))
