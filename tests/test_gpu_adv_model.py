"""GPU (-m gpu): the caller-side steps fused into dp_purify (bilinear resize, range maps, classifier normalisation:
SURVEY.md section 8f-2) vs their torch restatement (oracle/prepost.py = F.interpolate, which is what the reference calls),
the SDE_Adv_Model drop-in class on the engine, and nn.DataParallel over two GPUs (eval_sde_adv.py:227-229)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import ncsnpp as O, prepost as PP, weights

pytestmark = pytest.mark.gpu


def _tiny():
    cfg = O.tiny_cfg(64, (1, 2, 2), 1, (16,), 32)
    return cfg, weights.make_state_dict(O.param_shapes(cfg), seed=2)


def _config(cfg, device):
    c = SimpleNamespace(data=SimpleNamespace(dataset="CIFAR10", image_size=cfg.image_size, num_channels=3),
                        model=SimpleNamespace(name="ncsnpp", resblock_type="biggan", fir=False, skip_rescale=True,
                                              progressive="none", progressive_input="none",
                                              embedding_type="positional", conditional=True, nonlinearity="swish",
                                              nf=cfg.nf, ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                                              attn_resolutions=list(cfg.attn_resolutions)))
    c.device = device
    return c


class _Wrapper(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        conv = torch.nn.Conv2d(3, 5, 3, padding=1)
        with torch.no_grad():
            conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.2)
            conv.bias.zero_()
        self.resnet = torch.nn.Sequential(conv, torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten())
        self.mu = torch.Tensor(PP.IMAGENET_MU).float().view(3, 1, 1)
        self.sigma = torch.Tensor(PP.IMAGENET_SIGMA).float().view(3, 1, 1)

    def forward(self, x):
        return self.resnet((x - self.mu.to(x.device)) / self.sigma.to(x.device))


@pytest.mark.parametrize("in_hw,out_hw,norm", [((28, 28), (28, 28), True), ((32, 32), None, False), ((40, 24), (20, 36), True)])
def test_fused_pre_post_kernels_match_torch(in_hw, out_hw, norm):
    """Identity loop (coefficients (1, 0, 0), no forward-diffusion noise): out == post(pre(x)) of the torch restatement,
    i.e. F.interpolate(bilinear, align_corners=False) both ways (the 224 <-> 256 ratio 7/8 of the ImageNet path, and an
    anisotropic case), the range maps and the classifier normalisation."""
    from diffpure_b200 import lowering_ncsnpp as L
    from diffpure_b200.engine import Engine
    cfg, sd = _tiny()
    lcfg = SimpleNamespace(image_size=32, num_channels=3, nf=cfg.nf, ch_mult=cfg.ch_mult,
                           num_res_blocks=cfg.num_res_blocks, attn_resolutions=cfg.attn_resolutions)
    eng = Engine(L.lower(lcfg, sd, 3), device=0)
    g = torch.Generator().manual_seed(1)
    x01 = torch.rand(3, 3, *in_hw, generator=g)
    nrm = (PP.IMAGENET_MU, PP.IMAGENET_SIGMA) if norm else None
    cond = np.full(2, 10.0, np.float32)
    coef = np.tile(np.array([[1.0, 0.0, 0.0]], np.float32), (2, 1))
    out = eng.purify(x01.cuda(), cond, coef, 1.0, 0.0, in_unit_range=True, out_hw=out_hw, out_unit_range=True,
                     out_norm=nrm).cpu()
    eng.close()
    want = PP.post(PP.pre(x01, (32, 32)), out_hw, nrm)
    assert out.shape == want.shape
    assert (out - want).abs().max().item() < 2e-6, (out - want).abs().max().item()


def test_sde_adv_model_on_the_engine():
    """The drop-in class: fused forward (no grad) == the torch-op composition around the same engine call; a requires_grad
    input takes the differentiable path."""
    from diffpure_b200.adv_model import SDE_Adv_Model
    cfg, sd = _tiny()
    dev = torch.device("cuda:0")
    args = SimpleNamespace(t=4, rand_t=False, t_delta=3, use_bm=False, score_type="score_sde", sample_step=1,
                           log_dir="/tmp/dp_test_logs", save_images=False, diffusion_type="sde", domain="cifar10",
                           classifier_name="stub")
    clf = _Wrapper()
    m = SDE_Adv_Model(args, _config(cfg, dev), classifier=clf, state_dict=sd).eval()
    g = torch.Generator().manual_seed(3)
    x = torch.rand(4, 3, 32, 32, generator=g).cuda()
    np.random.seed(11)
    with torch.no_grad():
        logits = m(x)
    assert logits.shape == (4, 5) and torch.isfinite(logits).all() and int(m.counter.item()) == 1
    # the same call by hand: engine generator for both noises (seed drawn from NumPy as the runner does)
    np.random.seed(11)
    seed = int(np.random.randint(0, 2 ** 31 - 1))
    from diffpure_b200 import schedule
    cond, coef = schedule.vpsde_tables(4)
    sx, se = schedule.vpsde_forward_scales(4)
    eng = m.runner.model.engine_for(4, dev)
    xr = eng.purify(PP.pre(x.cpu()).cuda(), cond, coef, sx, se, seed=seed)
    want = clf.to(dev)((xr + 1) * 0.5)
    assert torch.allclose(logits, want, atol=1e-4, rtol=1e-4), (logits - want).abs().max().item()
    xg = x.clone().requires_grad_(True)
    m(xg).sum().backward()
    assert xg.grad is not None and torch.isfinite(xg.grad).all() and xg.grad.abs().max().item() > 0
    m.runner.model.release()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_data_parallel_two_gpus_matches_single_gpu_shards():
    """nn.DataParallel over the runner, as eval_sde_adv.py:227-229 wraps SDE_Adv_Model: each replica purifies its chunk on
    its own GPU with its own engine and weight blob; results equal the single-GPU runs of the two halves bit for bit."""
    from diffpure_b200.runners.diffpure_sde import RevGuidedDiffusion
    cfg, sd = _tiny()
    args = SimpleNamespace(t=4, rand_t=False, t_delta=3, use_bm=False, score_type="score_sde", sample_step=1,
                           log_dir="/tmp/dp_test_logs", save_images=False)

    class Wrap(torch.nn.Module):
        def __init__(self, runner):
            super().__init__()
            self.runner = runner

        def forward(self, x):
            return self.runner.image_editing_sample(x, bs_id=3, tag="dp", seed=21)

    runner = RevGuidedDiffusion(args, _config(cfg, torch.device("cuda")), device=torch.device("cuda"), state_dict=sd)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(8, 3, 32, 32, generator=g) * 2 - 1)
    torch.manual_seed(0)
    with torch.no_grad():
        torch.manual_seed(0)
        lo = runner.image_editing_sample(x[:4].cuda(0), bs_id=3, tag="s", seed=21, init_noise=torch.zeros(4, 3, 32, 32))
        hi = runner.image_editing_sample(x[4:].cuda(0), bs_id=3, tag="s", seed=21, init_noise=torch.zeros(4, 3, 32, 32))

    class WrapZ(Wrap):
        def forward(self, x):
            return self.runner.image_editing_sample(x, bs_id=3, tag="dp", seed=21, init_noise=torch.zeros_like(x))

    dp = torch.nn.DataParallel(WrapZ(runner), device_ids=[0, 1])
    with torch.no_grad():
        out = dp(x.cuda(0))
    assert out.shape == (8, 3, 32, 32) and out.device.index == 0
    assert torch.equal(out[:4], lo) and torch.equal(out[4:].cpu(), hi.cpu())
    keys = sorted(k for k in runner.model._engines)
    assert {k[1] for k in keys} == {0, 1}, keys          # one engine (and weight blob) per GPU
    runner.model.release()
