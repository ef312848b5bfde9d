"""CPU: the oracle of the next scope row (input gradients for white-box attacks): the
kernel-shaped backward restatement oracle/ncsnpp_vjp.py vs torch.autograd on the (reference-identical) oracle forward."""
import torch

from oracle import ncsnpp as O, ncsnpp_vjp as V, sde as OS, weights


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_primitives_match_autograd():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 16, 8, 8, generator=g, requires_grad=True)
    w = torch.randn(24, 16, 3, 3, generator=g) * 0.1
    go = torch.randn(2, 24, 8, 8, generator=g)
    (ref,) = torch.autograd.grad(torch.nn.functional.conv2d(x, w, padding=1), x, go)
    assert _rel(V.conv_dgrad(go, w), ref) < 1e-5
    gam, bet = torch.randn(16, generator=g), torch.randn(16, generator=g)
    gy = torch.randn(2, 16, 8, 8, generator=g)
    for silu in (True, False):
        y = O._gn(x, gam, bet)
        y = torch.nn.functional.silu(y) if silu else y
        (ref,) = torch.autograd.grad(y, x, gy)
        assert _rel(V.gn_silu_vjp(x.detach(), gam, gy, silu, bet), ref) < 1e-4
    (ref,) = torch.autograd.grad(O._up2(x), x, torch.ones(2, 16, 16, 16))
    assert torch.allclose(V.up2_vjp(torch.ones(2, 16, 16, 16)), ref)
    gd = torch.randn(2, 16, 4, 4, generator=g)
    (ref,) = torch.autograd.grad(O._down2(x), x, gd)
    assert torch.allclose(V.down2_vjp(gd), ref, atol=1e-6)


def test_network_vjp_matches_autograd():
    for cfg, seed in ((O.tiny_cfg(64, (1, 2), 1, (8,), 16), 1), (O.tiny_cfg(64, (1, 2, 2), 2, (16,), 32), 2)):
        sd = weights.make_state_dict(O.param_shapes(cfg), seed=seed)
        g = torch.Generator().manual_seed(seed)
        S = cfg.image_size
        x = (torch.rand(2, 3, S, S, generator=g) * 2 - 1).requires_grad_(True)
        t = torch.tensor([37.0, 512.0])
        go = torch.randn(2, 3, S, S, generator=g)
        y = O.forward(cfg, sd, x, t)
        (ref,) = torch.autograd.grad(y, x, go)
        with torch.no_grad():
            y2, _ = V.forward_with_tape(cfg, sd, x.detach(), t)
            got = V.vjp(cfg, sd, x.detach(), t, go)
        assert torch.allclose(y2, y.detach(), atol=1e-5)
        assert _rel(got, ref) < 1e-4, _rel(got, ref)


def test_loop_vjp_matches_autograd_through_the_euler_loop():
    cfg = O.tiny_cfg(64, (1, 2), 1, (8,), 16)
    sd = weights.make_state_dict(O.param_shapes(cfg), seed=3)
    g = torch.Generator().manual_seed(3)
    t_star = 4
    n = OS.num_steps(t_star)
    x0 = (torch.rand(2, 3, 16, 16, generator=g) * 2 - 1).requires_grad_(True)
    e0 = torch.randn(2, 3, 16, 16, generator=g)
    z = torch.randn(n, 2, 3, 16, 16, generator=g)
    go = torch.randn(2, 3, 16, 16, generator=g)
    out = OS.purify_sde(lambda xx, tt: O.forward(cfg, sd, xx, tt), x0, t_star, e0, z)
    (ref,) = torch.autograd.grad(out, x0, go)
    got, out2 = V.purify_sde_vjp(cfg, sd, x0.detach(), t_star, e0, z, go)
    assert torch.allclose(out2, out.detach(), atol=1e-4)
    assert _rel(got, ref) < 1e-3, _rel(got, ref)
