"""CPU: the product's per-step scalar tables agree with the oracle's step-by-step arithmetic."""
import numpy as np
import torch

from diffpure_b200 import schedule
from oracle import sde as OS


def test_vpsde_tables_reproduce_oracle_update():
    for t_star in (4, 100, 150):
        cond, coef = schedule.vpsde_tables(t_star, "score_sde")
        grid = OS.time_grid(t_star)
        assert len(cond) == len(grid) - 1
        g = torch.Generator().manual_seed(0)
        x = torch.randn(2, 3, 8, 8, generator=g)
        eps = torch.randn(2, 3, 8, 8, generator=g)
        z = torch.randn(2, 3, 8, 8, generator=g)
        for k in (0, len(cond) // 2, len(cond) - 1):
            t, tn = grid[k], grid[k + 1]
            h = tn - t
            labels = []
            f = OS.rev_vpsde_f(lambda xx, tt: (labels.append(tt), eps)[1], "score_sde", t, x)
            ref = x + f * h + OS.rev_vpsde_g(t, 2)[:, None, None, None] * (z * torch.sqrt(h))
            mine = coef[k, 0] * x + coef[k, 1] * eps + coef[k, 2] * z
            assert (ref - mine).abs().max().item() < 2e-6
            assert abs(float(labels[0][0]) - float(cond[k])) < 1e-4


def test_guided_timesteps_floor_in_fp32():
    """(s*1000).long() on the fp32 grid: t*=150 starts at 149, contains a duplicate, ends at 1."""
    cond, _ = schedule.vpsde_tables(150, "guided_diffusion")
    c = cond.astype(np.int64)
    assert c[0] == 149 and c[-1] == 1 and len(c) == 150
    assert (np.diff(c) <= 0).all() and (np.diff(c) == 0).sum() >= 1


def test_forward_scales():
    sx, se = schedule.vpsde_forward_scales(100)
    assert abs(sx * sx + se * se - 1.0) < 1e-6 and 0.9 < sx < 1.0
