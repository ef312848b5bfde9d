"""CPU: diffpure_b200.adv_model.SDE_Adv_Model (reference interface, eval_sde_adv.py:34-93) -- the torch-op path taken for
CPU tensors / gradient requests equals the oracle's pre / post restatement around the runner call; counters and tags
behave like the reference's. (The fused GPU path is held to the same restatement by tests/test_gpu_adv_model.py.)"""
from types import SimpleNamespace

import numpy as np
import torch

from oracle import prepost as PP
from test_runners_cpu import FakeEngine, StubNet, _cifar_config, _cifar_sd


class _Wrapper(torch.nn.Module):
    """Shape of the reference's classifier wrappers (utils.py:144-153): normalise, then the network."""

    def __init__(self):
        super().__init__()
        self.resnet = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.AdaptiveAvgPool2d(1),
                                          torch.nn.Flatten())
        self.mu = torch.Tensor(PP.IMAGENET_MU).float().view(3, 1, 1)
        self.sigma = torch.Tensor(PP.IMAGENET_SIGMA).float().view(3, 1, 1)

    def forward(self, x):
        return self.resnet((x - self.mu.to(x.device)) / self.sigma.to(x.device))


def test_unfused_path_matches_the_reference_composition(capsys):
    from diffpure_b200.adv_model import SDE_Adv_Model
    args = SimpleNamespace(t=5, rand_t=False, t_delta=3, use_bm=False, score_type="score_sde", sample_step=1,
                           log_dir="/tmp/dp_cpu_logs", save_images=False, diffusion_type="sde", domain="imagenet",
                           classifier_name="stub")
    config = _cifar_config()
    config.device = torch.device("cpu")
    torch.manual_seed(0)
    clf = _Wrapper()
    m = SDE_Adv_Model(args, config, classifier=clf, state_dict=_cifar_sd())
    net = StubNet(3)
    m.runner.model = net
    m.runner.rev_vpsde.model = net
    m.set_tag("tg")
    x = torch.rand(2, 3, 14, 14)                      # 'imagenet' domain: resized to 256 by the reference; the stub net is
    import torch.nn.functional as F                  # size-agnostic, so the 256-pixel round trip runs on the CPU in a moment
    np.random.seed(5)
    torch.manual_seed(9)
    out = m(x)
    assert "diffusion times: 0" in capsys.readouterr().out
    # manual composition with the same RNG streams
    np.random.seed(5)
    torch.manual_seed(9)
    x256 = PP.pre(x, (256, 256))
    e = torch.randn_like(x256)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    from diffpure_b200 import schedule
    cond, coef = schedule.vpsde_tables(5)
    sx, se = schedule.vpsde_forward_scales(5)
    xr = FakeEngine(net).purify(x256, cond, coef, sx, se, init_noise=e, seed=seed)
    want = clf.resnet(PP.post(xr, (224, 224), (PP.IMAGENET_MU, PP.IMAGENET_SIGMA)))
    assert torch.allclose(out, want, atol=1e-5)
    assert F.interpolate(xr, size=(224, 224), mode="bilinear", align_corners=False).shape[-1] == 224
    assert int(m.counter.item()) == 1
    m.reset_counter()
    assert int(m.counter.item()) == 0 and m._count == 0
    # gradients flow through the torch-op path (white-box attacks)
    xg = torch.rand(2, 3, 14, 14, requires_grad=True)
    m(xg).sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all() and xg.grad.abs().max() > 0


def test_bpda_variant_modes(capsys):
    """eval_sde_adv_bpda.py:53-117: `resnet`, purify() -> [0,1] images, forward(x, mode) with the three modes."""
    import pytest
    from diffpure_b200.adv_model import SDE_Adv_Model_BPDA
    args = SimpleNamespace(t=5, rand_t=False, t_delta=3, use_bm=False, score_type="score_sde", sample_step=1,
                           log_dir="/tmp/dp_cpu_logs", save_images=False, diffusion_type="sde", domain="cifar10",
                           classifier_name="stub")
    config = _cifar_config()
    config.device = torch.device("cpu")
    torch.manual_seed(0)
    clf = _Wrapper()
    m = SDE_Adv_Model_BPDA(args, config, classifier=clf, state_dict=_cifar_sd())
    assert m.resnet is clf and "resnet.resnet.0.weight" not in m.state_dict()      # an alias, not a second registration
    net = StubNet(3)
    m.runner.model = net
    m.runner.rev_vpsde.model = net
    x = torch.rand(2, 3, 16, 16)

    def run(mode):
        np.random.seed(5)
        torch.manual_seed(9)
        return m(x, mode=mode)

    pur = run("purify")
    assert pur.shape == x.shape and "diffusion times: 0" in capsys.readouterr().out
    # the same composition by hand (cifar10 domain: no resize)
    np.random.seed(5)
    torch.manual_seed(9)
    from diffpure_b200 import schedule
    x2 = (x - 0.5) * 2
    e = torch.randn_like(x2)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    cond, coef = schedule.vpsde_tables(5)
    xr = FakeEngine(net).purify(x2, cond, coef, *schedule.vpsde_forward_scales(5), init_noise=e, seed=seed)
    assert torch.allclose(pur, (xr + 1) * 0.5, atol=1e-6)
    assert torch.allclose(run("purify_and_classify"), clf(pur), atol=1e-6)
    assert torch.allclose(m(pur, mode="classify"), clf(pur))
    assert int(m.counter.item()) == 2                     # 'classify' does not count as a diffusion call (L107-108)
    with pytest.raises(NotImplementedError):
        m(x, mode="nope")
    args.diffusion_type = "ode"                           # the BPDA script knows ddpm / sde / celebahq-ddpm only
    with pytest.raises(NotImplementedError):
        SDE_Adv_Model_BPDA(args, config, classifier=clf, state_dict=_cifar_sd())
