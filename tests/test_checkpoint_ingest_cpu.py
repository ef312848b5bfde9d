"""CPU: checkpoint ingest. The score_sde checkpoint layout -- DataParallel 'module.' keys, the
'sigmas' buffer, EMA shadow parameters that replace the weights -- is loaded by diffpure_b200 exactly as the reference's
restore_checkpoint + ExponentialMovingAverage.copy_to do (runners/diffpure_sde.py:42-47,175-182; score_sde/models/ema.py:61-72)."""
import os

import pytest
import torch

REF = os.path.isdir("/root/reference/score_sde")


def _fake_checkpoint(tmp_path):
    """A checkpoint in the reference's on-disk layout, written by hand (no reference code needed)."""
    from diffpure_b200 import lowering_ncsnpp as L, synthetic
    from types import SimpleNamespace
    cfg = SimpleNamespace(image_size=16, num_channels=3, nf=64, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(8,))
    shapes = L.param_shapes(cfg)
    model_sd = {"module." + k: v for k, v in synthetic.random_state_dict(shapes, seed=3).items()}
    model_sd["module.sigmas"] = torch.linspace(50.0, 0.01, 1000)
    ema_sd = synthetic.random_state_dict(shapes, seed=4)
    state = {"optimizer": {}, "model": model_sd, "step": 7,
             "ema": {"decay": 0.9999, "num_updates": 7, "shadow_params": [ema_sd[k] for k in shapes]}}
    path = os.path.join(tmp_path, "checkpoint_8.pth")
    torch.save(state, path)
    return path, ema_sd


def test_score_sde_checkpoint_layout(tmp_path):
    from diffpure_b200.runners.diffpure_sde import _load_score_sde_state
    path, ema_sd = _fake_checkpoint(str(tmp_path))
    sd = _load_score_sde_state(path)
    assert "sigmas" in sd and not any(k.startswith("module.") for k in sd)
    for k, v in ema_sd.items():
        assert torch.equal(sd[k], v), k                  # EMA shadow replaced every weight, in parameter order


@pytest.mark.skipif(not REF, reason="reference tree not present")
@pytest.mark.parametrize("wrap", [False, True])
def test_score_sde_checkpoint_matches_reference_loader(tmp_path, wrap):
    """Same file through the reference's own classes (NCSNpp + DataParallel + optimizer + EMA) and through ours."""
    from oracle import ref_import
    ref_import.install()
    over = dict(nf=64, ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[8], **{"data.image_size": 16})
    model, cfg = ref_import.build_ncsnpp(over)
    if wrap:                                              # checkpoints written from a DataParallel model carry 'module.' keys
        model = torch.nn.DataParallel(model)
    from score_sde.models.ema import ExponentialMovingAverage
    from score_sde.losses import get_optimizer
    from runners.diffpure_sde import restore_checkpoint
    torch.manual_seed(0)
    for p in model.parameters():                          # zero-initialised tensors would hide ordering mistakes
        p.data.normal_()
    ema = ExponentialMovingAverage(model.parameters(), decay=0.5)
    for p in model.parameters():
        p.data.add_(torch.randn_like(p))
    ema.update(model.parameters())                        # shadow != weights
    opt = get_optimizer(cfg, model.parameters())
    path = os.path.join(str(tmp_path), "checkpoint_8.pth")
    torch.save({"optimizer": opt.state_dict(), "model": model.state_dict(), "ema": ema.state_dict(), "step": 3}, path)

    model2, cfg2 = ref_import.build_ncsnpp(over)
    if wrap:
        model2 = torch.nn.DataParallel(model2)
    ema2 = ExponentialMovingAverage(model2.parameters(), decay=cfg2.model.ema_rate)
    state = dict(step=0, optimizer=get_optimizer(cfg2, model2.parameters()), model=model2, ema=ema2)
    restore_checkpoint(path, state, "cpu")                # runners/diffpure_sde.py:42-47
    ema2.copy_to(model2.parameters())                     # L182
    want = (model2.module if wrap else model2).state_dict()

    from diffpure_b200.runners.diffpure_sde import _load_score_sde_state
    got = _load_score_sde_state(path)
    assert set(got) == set(want)
    for k, v in want.items():
        assert torch.equal(got[k], v), k


def _ref_yaml(name):
    """The shipped yaml files' model sections, restated (configs/*.yml) -- what the runners receive as `config`."""
    from types import SimpleNamespace as NS
    if name == "cifar10":
        return NS(data=NS(dataset="CIFAR10", image_size=32, num_channels=3),
                  model=NS(name="ncsnpp", resblock_type="biggan", fir=False, skip_rescale=True, progressive="none",
                           progressive_input="none", embedding_type="positional", conditional=True, nonlinearity="swish",
                           nf=128, ch_mult=[1, 2, 2, 2], num_res_blocks=8, attn_resolutions=[16]))
    if name == "imagenet":
        return NS(data=NS(dataset="ImageNet"),
                  model=NS(attention_resolutions="32,16,8", class_cond=False, diffusion_steps=1000, rescale_timesteps=True,
                           timestep_respacing="1000", image_size=256, learn_sigma=True, noise_schedule="linear",
                           num_channels=256, num_head_channels=64, num_res_blocks=2, resblock_updown=True, use_fp16=True,
                           use_scale_shift_norm=True))
    return NS(data=NS(dataset="CelebA_HQ", image_size=256),
              model=NS(ch=128, out_ch=3, ch_mult=[1, 1, 2, 2, 4, 4], num_res_blocks=2, attn_resolutions=[16], in_channels=3,
                       resamp_with_conv=True, var_type="fixedsmall"))


def test_real_checkpoint_key_sets():
    """The three real checkpoints are loaded by the reference with strict load_state_dict, so their key sets are the
    state_dict() of the reference modules built from the shipped configs (fixture: oracle/make_golden.py
    --checkpoint-keys). The lowerings must consume exactly those names and shapes (plus nothing else)."""
    import json
    from diffpure_b200 import lowering_adm as LA, lowering_ddpm as LD, lowering_ncsnpp as LN
    with open(os.path.join(os.path.dirname(__file__), "golden", "checkpoint_keys.json")) as f:
        keys = json.load(f)
    want = {k: {n: tuple(s) for n, s in v} for k, v in keys.items()}
    ck = want["score_sde/checkpoint_8.pth:model (configs/cifar10.yml)"]
    got = LN.param_shapes(LN.cfg_from_reference(_ref_yaml("cifar10")))
    assert ck.pop("sigmas") == (1000,)                      # the one buffer the score network never reads (ncsnpp.py:59)
    assert {k: tuple(v) for k, v in got.items()} == ck and list(got) == list(ck)   # same order: the EMA shadow list is positional
    ck = want["guided_diffusion/256x256_diffusion_uncond.pt (configs/imagenet.yml)"]
    got = LA.param_shapes(LA.cfg_from_reference(_ref_yaml("imagenet")))
    assert {k: tuple(v) for k, v in got.items()} == ck
    ck = want["celeba_hq.ckpt (configs/celeba.yml)"]
    got = LD.param_shapes(LD.cfg_from_reference(_ref_yaml("celeba")))
    assert {k: tuple(v) for k, v in got.items()} == ck


@pytest.mark.skipif(not REF, reason="reference tree not present")
def test_reference_module_state_dicts_load_directly():
    """state_dict() of the reference's own (reduced) ADM -- after convert_to_fp16, as GuidedDiffusion holds it
    (diffpure_guided.py:31-35) -- and CelebA modules go straight into the runners: names, shapes and dtypes are accepted
    and the lowered programs reproduce the reference modules through the CPU interpreter."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from oracle import ref_import
    from program_interp import Interp
    from diffpure_b200 import lowering_adm as LA, lowering_ddpm as LD
    from types import SimpleNamespace as NS
    torch.manual_seed(0)
    m, _, mc = ref_import.build_adm(num_channels=64, image_size=64, num_res_blocks=1, use_fp16=True)
    with torch.no_grad():
        for p_ in m.parameters():                            # zero-initialised layers would make the check vacuous
            if p_.abs().max() == 0:
                p_.copy_((torch.randn_like(p_.float()) * 0.02).to(p_.dtype))
    sd = m.state_dict()
    assert any(v.dtype == torch.float16 for v in sd.values())
    cfg = LA.cfg_from_reference(NS(model=NS(**mc)))
    assert set(LA.param_shapes(cfg)) == set(sd)
    x = torch.rand(1, 3, 64, 64) * 2 - 1
    t = torch.tensor([17])
    with torch.no_grad():
        y = m(x, t).float()
    sd32 = {k: v.float() for k, v in sd.items()}
    got = Interp(LA.lower(cfg, sd32, 1), emulate_bf16=False).run(x, t.float())
    assert ((got - y).norm() / y.norm()).item() < 2e-2       # the reference ran its torso in fp16
    mc_, ccfg = ref_import.build_celeba({"ch": 64, "ch_mult": [1, 2, 2], "num_res_blocks": 1, "attn_resolutions": [16],
                                         "data.image_size": 32})
    sdc = mc_.state_dict()
    lcfg = LD.cfg_from_reference(ccfg)
    assert set(LD.param_shapes(lcfg)) == set(sdc)
    xc = torch.rand(1, 3, 32, 32) * 2 - 1
    with torch.no_grad():
        yc = mc_(xc, torch.tensor([9]))
    gotc = Interp(LD.lower(lcfg, sdc, 1), emulate_bf16=False).run(xc, torch.tensor([9.0]))
    assert ((gotc - yc).norm() / yc.norm()).item() < 1e-4
