"""Shared state of the full-size -m gpu cases: the seeded random-init weights of the real shapes (5.7 s to draw for the 552.8 M
parameter ImageNet network), one packed weight blob per model on cuda:0 (engines of every batch size adopt it), and the
precomputed oracle results (tests/golden/fullsize_oracle.npz, oracle/make_fullsize_golden.py)."""
import functools
import os

import numpy as np
import torch

import golden_inputs as GI

G = os.path.join(os.path.dirname(__file__), "golden")


@functools.lru_cache(maxsize=None)
def oracle_results():
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "fullsize_oracle.npz")).items()}


def _lowering(which):
    if which == "adm":
        from diffpure_b200 import lowering_adm as L
        return L, L.imagenet_cfg()
    if which == "celeba":
        from diffpure_b200 import lowering_ddpm as L
        return L, L.celeba_cfg()
    raise KeyError(which)


@functools.lru_cache(maxsize=None)
def state_dict(which):
    """Seeded weights; checked against the fingerprint stored with the oracle results (the fixture is void otherwise)."""
    if which == "cifar":
        from oracle import ncsnpp as O, weights
        sd = weights.make_state_dict(O.param_shapes(O.CIFAR10_CFG), seed=0)
    else:
        from diffpure_b200 import synthetic
        L, cfg = _lowering(which)
        sd = synthetic.random_state_dict(L.param_shapes(cfg), seed=0)
    want = oracle_results()[f"fp_{which}"]
    got = GI.weights_fingerprint(sd)
    assert torch.allclose(got, want, rtol=1e-9, atol=0), \
        f"seeded {which} weights differ from the ones tests/golden/fullsize_oracle.npz was computed with: {got} vs {want}"
    return sd


@functools.lru_cache(maxsize=None)
def _blob(which):
    from diffpure_b200.engine import WeightBlob
    L, cfg = _lowering(which)
    return WeightBlob(L.lower(cfg, state_dict(which), 1), 0)


def engine(which, B):
    """Forward engine of a 256x256 model at batch B on cuda:0, sharing the model's packed weight blob."""
    from diffpure_b200.engine import Engine
    L, cfg = _lowering(which)
    return Engine(L.lower(cfg, state_dict(which), B), device=0, blob=_blob(which))


def sparse(t):
    return GI.at_pixels(t, GI.sparse_pixels())
