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
