"""GPU (-m gpu): the input gradient of the score network and of the purification loop through the C ABI
(dp_unet_vjp + the runners' torch.autograd.Function) vs oracle/ncsnpp_vjp.py (CPU fp32, held to torch.autograd).

Stated tolerance for gradients (bf16 tensor-core operands in both passes, fp32 accumulation, fp32 gradient stream):
rel-L2 <= 2e-2 of the oracle gradient, the same bound as one forward evaluation -- rounding the frozen weights to bf16
alone moves the input gradient of the reduced networks by ~1.4e-2 (tests/test_vjp_lowering_cpu.py). Measured on B200
(profiles/r02_vjp_parity.log): 1.57e-2 / 1.83e-2 on the reduced networks, 1.21e-2 on the full CIFAR-10 model,
5.4e-4 for the gradient through a 3-step loop."""
from types import SimpleNamespace

import pytest
import torch

from golden_inputs import cifar_vjp_inputs, fullsize_adm_vjp_inputs
from oracle import ncsnpp as O, ncsnpp_vjp as V, weights

pytestmark = pytest.mark.gpu
TOL_VJP = 2e-2


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def vjp_engine(cfg, sd, B):
    from diffpure_b200 import lowering_ncsnpp as L
    from diffpure_b200.engine import Engine
    lcfg = SimpleNamespace(image_size=cfg.image_size, num_channels=3, nf=cfg.nf, ch_mult=cfg.ch_mult,
                           num_res_blocks=cfg.num_res_blocks, attn_resolutions=cfg.attn_resolutions)
    return Engine(L.lower_vjp(lcfg, sd, B), device=0)


@pytest.mark.parametrize("name,cfg,seed", [
    ("tiny-small-attn", O.tiny_cfg(64, (1, 2), 1, (8,), 16), 1),
    ("tiny-tc-attn", O.tiny_cfg(64, (1, 2, 2), 2, (16,), 32), 2),
    ("cifar10-full", O.CIFAR10_CFG, 0),
])
def test_unet_vjp_vs_oracle(name, cfg, seed):
    S, B = cfg.image_size, 2
    if name == "cifar10-full":       # the oracle gradient of the 106 M parameter model is precomputed (fullsize_oracle.npz)
        import fullsize as F
        sd = F.state_dict("cifar")
        x, t, go = cifar_vjp_inputs(seed)
        ref = F.oracle_results()["cifar_vjp"]
    else:
        sd = weights.make_state_dict(O.param_shapes(cfg), seed=seed)
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(B, 3, S, S, generator=g) * 2 - 1
        t = torch.tensor([37.0, 512.0])
        go = torch.randn(B, 3, S, S, generator=g)
        with torch.no_grad():
            ref = V.vjp(cfg, sd, x, t, go)
    eng = vjp_engine(cfg, sd, B)
    got = eng.unet_vjp(x.cuda(), t.cuda(), go.cuda()).cpu()
    got2 = eng.unet_vjp(x.cuda(), t.cuda(), go.cuda()).cpu()
    eng.close()
    r = rel(got, ref)
    print(f"unet vjp {name}: rel-L2 vs the oracle gradient = {r:.3e}")
    assert torch.equal(got, got2)                       # deterministic: no atomics in the backward kernels either
    assert r < TOL_VJP, r


def _cifar_like_config(cfg):
    return SimpleNamespace(data=SimpleNamespace(dataset="CIFAR10", image_size=cfg.image_size, num_channels=3),
                           model=SimpleNamespace(name="ncsnpp", resblock_type="biggan", fir=False, skip_rescale=True,
                                                 progressive="none", progressive_input="none",
                                                 embedding_type="positional", conditional=True, nonlinearity="swish",
                                                 nf=cfg.nf, ch_mult=list(cfg.ch_mult),
                                                 num_res_blocks=cfg.num_res_blocks,
                                                 attn_resolutions=list(cfg.attn_resolutions)))


def test_runner_gradient_through_a_three_step_loop():
    """RevGuidedDiffusion.image_editing_sample with a requires_grad input (the white-box attack path of
    eval_sde_adv.py:126-128): d<w, purified>/dx0 over K = 3 Euler-Maruyama steps vs the oracle's discrete adjoint."""
    from diffpure_b200.runners.diffpure_sde import RevGuidedDiffusion
    cfg = O.tiny_cfg(64, (1, 2, 2), 1, (16,), 32)
    sd = weights.make_state_dict(O.param_shapes(cfg), seed=2)
    t_star = 3
    args = SimpleNamespace(t=t_star, rand_t=False, t_delta=15, use_bm=False, score_type="score_sde", sample_step=1,
                           log_dir="/tmp/dp_test_logs", save_images=False)
    runner = RevGuidedDiffusion(args, _cifar_like_config(cfg), device=torch.device("cuda:0"), state_dict=sd)
    g = torch.Generator().manual_seed(4)
    x0 = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    e0 = torch.randn(2, 3, 32, 32, generator=g)
    z = torch.randn(3, 2, 3, 32, 32, generator=g)
    w = torch.randn(2, 3, 32, 32, generator=g)
    ref, out_ref = V.purify_sde_vjp(cfg, sd, x0, t_star, e0, z, w)
    xg = x0.cuda().requires_grad_(True)
    out = runner.image_editing_sample(xg, bs_id=5, tag="g", init_noise=e0.cuda(), step_noise=z.cuda())
    assert out.requires_grad
    (gx,) = torch.autograd.grad((out * w.cuda()).sum(), xg)
    r_out, r_g = rel(out.detach().cpu(), out_ref), rel(gx.cpu(), ref)
    print(f"K=3 loop: state rel-L2 {r_out:.3e}, input-gradient rel-L2 {r_g:.3e}")
    assert r_out < 5e-3 and r_g < TOL_VJP, (r_out, r_g)
    # the no-grad path is unchanged and agrees with the differentiable one
    with torch.no_grad():
        out2 = runner.image_editing_sample(x0.cuda(), bs_id=5, tag="g", init_noise=e0.cuda(), step_noise=z.cuda())
    assert torch.equal(out2, out.detach())
    runner.model.release()


def _adm_tiny():
    from oracle import adm as A
    acfg = A.tiny_cfg(64, 64, (1, 2, 3, 4), 1, (32, 16, 8))     # attention at T = 1024, 256 and 64; heads of 64 channels
    return A, acfg, weights.make_state_dict(A.param_shapes(acfg), seed=5)


def test_adm_unet_vjp_vs_autograd():
    """guided_diffusion UNet (scale-shift norm, resblock up / down, multi-head attention): dp_unet_vjp against
    torch.autograd through the reference-pinned oracle forward, gradient wrt the eps half of the output."""
    from diffpure_b200 import lowering_adm as LA
    from diffpure_b200.engine import Engine
    A, acfg, sd = _adm_tiny()
    B = 2
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1).requires_grad_(True)
    t = torch.tensor([37.0, 512.0])
    go = torch.randn(B, 3, 64, 64, generator=g)
    (A.forward(acfg, sd, x, t)[:, :3] * go).sum().backward()
    ref = x.grad
    eng = Engine(LA.lower_vjp(acfg, sd, B), device=0)
    got = eng.unet_vjp(x.detach().cuda(), t.cuda(), go.cuda()).cpu()
    got2 = eng.unet_vjp(x.detach().cuda(), t.cuda(), go.cuda()).cpu()
    eng.close()
    r = rel(got, ref)
    print(f"unet vjp adm-tiny: rel-L2 vs autograd on the oracle = {r:.3e}")
    assert torch.equal(got, got2)
    assert r < TOL_VJP, r


def test_adm_runner_gradient_through_a_three_step_loop():
    """The ImageNet white-box path (run_in_rand_inf.sh: RevGuidedDiffusion, score_type guided_diffusion, gradient through
    sdeint_adjoint): d<w, purified>/dx0 over K = 3 Euler-Maruyama steps vs autograd through the oracle loop."""
    from diffpure_b200.runners.diffpure_sde import RevGuidedDiffusion
    from golden_inputs import ADM_TINY_REF_CONFIG
    from oracle import sde as OS
    A, acfg, sd = _adm_tiny()
    t_star = 3
    args = SimpleNamespace(t=t_star, rand_t=False, t_delta=15, use_bm=False, score_type="guided_diffusion", sample_step=1,
                           log_dir="/tmp/dp_test_logs", save_images=False)
    config = SimpleNamespace(data=SimpleNamespace(dataset="ImageNet"), model=SimpleNamespace(**ADM_TINY_REF_CONFIG))
    runner = RevGuidedDiffusion(args, config, device=torch.device("cuda:0"), state_dict=sd)
    g = torch.Generator().manual_seed(4)
    x0 = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    e0 = torch.randn(2, 3, 64, 64, generator=g)
    z = torch.randn(OS.num_steps(t_star), 2, 3, 64, 64, generator=g)
    w = torch.randn(2, 3, 64, 64, generator=g)
    xr = x0.clone().requires_grad_(True)
    out_ref = OS.purify_sde(lambda xx, tt: A.forward(acfg, sd, xx, tt), xr, t_star, e0, z, score_type="guided_diffusion")
    (ref,) = torch.autograd.grad((out_ref * w).sum(), xr)
    xg = x0.cuda().requires_grad_(True)
    out = runner.image_editing_sample(xg, bs_id=5, tag="g", init_noise=e0.cuda(), step_noise=z.cuda())
    assert out.requires_grad
    (gx,) = torch.autograd.grad((out * w.cuda()).sum(), xg)
    r_out, r_g = rel(out.detach().cpu(), out_ref.detach()), rel(gx.cpu(), ref)
    print(f"ADM K=3 loop: state rel-L2 {r_out:.3e}, input-gradient rel-L2 {r_g:.3e}")
    assert r_out < 5e-3 and r_g < TOL_VJP, (r_out, r_g)
    runner.model.release()


def test_gradient_request_fails_loudly_where_no_backward_program_exists():
    """The DDPM (CelebA-HQ, SDEdit) network has no input-gradient program: a requires_grad input must raise, never detach."""
    from diffpure_b200.runners._common import PurifyRunner
    r = PurifyRunner()
    r.model = SimpleNamespace(_lower_vjp=None)
    with pytest.raises(NotImplementedError):
        r._wants_grad(torch.zeros(1, 3, 8, 8, device="cuda").requires_grad_(True))
    assert r._wants_grad(torch.zeros(1, 3, 8, 8, device="cuda")) is False


def test_adm_fullsize_unet_vjp_vs_autograd():
    """The full ImageNet network (256x256, 552.8 M parameters, attention at T = 1024 / 256 / 64 with 4 - 16 heads): dp_unet_vjp
    at B = 1 against torch.autograd through the oracle forward (random-init weights of the real shapes; the autograd side is
    precomputed -- tests/golden/fullsize_oracle.npz -- it alone took minutes on the GPU box's host cores)."""
    import fullsize as F
    from diffpure_b200 import lowering_adm as LA
    from diffpure_b200.engine import Engine
    x, t, go = fullsize_adm_vjp_inputs()
    ref = F.oracle_results()["adm_vjp"]
    eng = Engine(LA.lower_vjp(LA.imagenet_cfg(), F.state_dict("adm"), 1), device=0)
    got = eng.unet_vjp(x.cuda(), t.cuda(), go.cuda()).cpu()
    eng.close()
    assert torch.isfinite(got).all()
    r = rel(F.sparse(got), ref)
    print(f"unet vjp adm-full (256x256, B=1): rel-L2 vs autograd on the oracle = {r:.3e}")
    assert r < TOL_VJP, r
