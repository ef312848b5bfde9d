"""CPU: the oracle restatements reproduce the golden vectors generated from the reference's own modules
(oracle/make_golden.py). Tolerance: fp32 round-off (1e-5 abs on O(1) values)."""
import os

import numpy as np
import pytest
import torch

from oracle import ncsnpp as O, sde as OS, weights

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return {k: torch.from_numpy(v) if v.ndim else v for k, v in np.load(os.path.join(G, name)).items()}


def test_ncsnpp_cifar10_eval_matches_reference():
    d = load("ncsnpp_cifar10_eval.npz")
    sd = weights.make_state_dict(O.param_shapes(O.CIFAR10_CFG), seed=int(d["seed"]))
    y = O.forward(O.CIFAR10_CFG, sd, d["x"], d["labels"])
    assert (y - d["y"]).abs().max().item() < 1e-5
    assert d["y"].abs().mean().item() > 0.1  # non-vacuous (zero-init layers are re-randomised)


@pytest.mark.parametrize("name,cfg", [
    ("tinyA", O.tiny_cfg(64, (1, 2), 1, (8,), 16)),
    ("tinyB", O.tiny_cfg(64, (1, 2, 2), 1, (16,), 32)),
])
def test_tiny_eval_and_loop_match_reference(name, cfg):
    d = load(f"ncsnpp_{name}.npz")
    sd = weights.make_state_dict(O.param_shapes(cfg), seed=int(d["seed"]))
    unet = lambda x, t: O.forward(cfg, sd, x, t)  # noqa: E731
    assert (unet(d["x"], d["labels"]) - d["y"]).abs().max().item() < 1e-5
    t_star = int(d["t_star"])
    xs = OS.forward_diffuse(d["x0"], d["e0"], t_star)
    grid = OS.time_grid(t_star)
    f0 = OS.rev_vpsde_f(unet, "score_sde", grid[0], xs)
    assert (f0 - d["f0"]).abs().max().item() < 1e-4 * max(1.0, d["f0"].abs().max().item())
    g0 = OS.rev_vpsde_g(grid[0], xs.shape[0])
    assert (g0 - d["g0"]).abs().max().item() < 1e-6
    out = OS.purify_sde(unet, d["x0"], t_star, d["e0"], d["z"])
    assert (out - d["loop_out"]).abs().max().item() < 2e-5


def test_time_grid_properties():
    """Appendix A.3 of SURVEY.md: t*=100 -> 100 steps, t*=150 -> 150 steps, last step shorter, fp32 accumulation."""
    for t_star, n in [(100, 100), (150, 150), (4, 4)]:
        g = OS.time_grid(t_star)
        assert len(g) - 1 == n
        h = g[1:] - g[:-1]
        assert h.dtype == torch.float32 and float(h[-1]) < float(h[0]) and abs(float(h[0]) - 1e-3) < 1e-6
        assert float(g[-1]) == float(torch.tensor(1 - 1e-5, dtype=torch.float32))
