"""CPU: the DDPM++ lowering (topology, operand offsets, weight packing, statistics plumbing) replayed by the
program interpreter reproduces the oracle to fp32 round-off; with bf16 operand emulation it stays inside the
tolerance stated for the tensor-core path."""
import pytest
import torch

from diffpure_b200 import lowering_ncsnpp as L
from oracle import ncsnpp as O, weights
from program_interp import Interp

CASES = [
    ("small-attn", O.tiny_cfg(64, (1, 2), 1, (8,), 16), 3),
    ("tc-attn", O.tiny_cfg(64, (1, 2, 2), 1, (16,), 32), 2),
    ("odd-batch", O.tiny_cfg(64, (1, 2), 2, (4,), 8), 5),
    ("fused-attn-block", O.tiny_cfg(128, (1, 2), 1, (16,), 32), 2),   # T = 256 tokens x C = 256: the one-kernel attention block
]


@pytest.mark.parametrize("name,cfg,B", CASES)
def test_lowering_matches_oracle(name, cfg, B):
    torch.manual_seed(0)
    sd = weights.make_state_dict(O.param_shapes(cfg), seed=1)
    x = torch.rand(B, 3, cfg.image_size, cfg.image_size) * 2 - 1
    t = torch.rand(B) * 999
    y = O.forward(cfg, sd, x, t)
    prog = L.lower(cfg, sd, B)
    y32 = Interp(prog, emulate_bf16=False).run(x, t)
    assert ((y32 - y).norm() / y.norm()).item() < 1e-5
    y16 = Interp(prog, emulate_bf16=True).run(x, t)
    assert ((y16 - y).norm() / y.norm()).item() < 2e-2
    kinds = [op.kind for op in prog.ops]
    if name == "fused-attn-block":
        assert kinds.count("attn_block") == 3 and not any(op.kind == "gemm" and op.args["softmax"] for op in prog.ops)
        unfused = L.lower(cfg, sd, B, fuse_attn=False)
        assert [op.kind for op in unfused.ops].count("attn_block") == 0 and len(unfused.ops) > len(prog.ops)
        yu = Interp(unfused, emulate_bf16=False).run(x, t)
        assert ((yu - y32).norm() / y.norm()).item() < 1e-5
    else:
        assert "attn_block" not in kinds


def test_param_shapes_match_oracle_and_count():
    a = L.param_shapes(L.cifar10_cfg())
    b = O.param_shapes(O.CIFAR10_CFG)
    assert list(a.items()) == [(k, tuple(v)) for k, v in b.items()]
    n = 0
    for v in a.values():
        m = 1
        for s in v:
            m *= s
        n += m
    assert n == 106632579  # DDPM++ parameter count


def test_program_structure_full_model():
    """Full CIFAR-10 model: 76 res-blocks + 10 attention blocks (534 engine launches vs ~1,093 ATen ops in the reference)."""
    cfg = L.cifar10_cfg()
    plan = L.module_plan(cfg)
    assert sum(1 for k, _ in plan if k == "res") == 76
    assert sum(1 for k, _ in plan if k == "attn") == 10


def test_liveness_pooled_activation_placement_never_overlaps_live_tensors():
    """Engine._place_activations (the activation allocator): with pooling on, two tensors whose live ranges
    [first use, last use] intersect never share bytes, and pooling shrinks the footprint."""
    from diffpure_b200.engine import Engine
    cfg = L.cifar10_cfg()
    from diffpure_b200 import synthetic
    prog = L.lower(cfg, synthetic.random_state_dict(L.param_shapes(cfg), seed=0), 2)

    def place(pool):
        e = Engine.__new__(Engine)                     # no CUDA library: only the placement logic runs
        e.program, e._ptr, e._loc, e.act_bytes = prog, {}, {}, 0
        top = [1 << 20]

        def alloc(nbytes):
            ptr = top[0]
            top[0] += nbytes
            return len(e._loc), ptr
        e._alloc = alloc
        e._place_activations(pool)
        return e

    pooled, flat = place(True), place(False)
    assert pooled.act_bytes < 0.5 * flat.act_bytes
    first, last = prog.first_use(), prog.last_use()
    spans = sorted((pooled._ptr[t.index], pooled._ptr[t.index] + t.nbytes, first[t.index], last[t.index], t.name)
                   for t in prog.tensors if t.init is None and t.index in first)
    for i, (a0, a1, f0, l0, n0) in enumerate(spans):
        for b0, b1, f1, l1, n1 in spans[i + 1:]:
            if b0 >= a1:
                break
            assert l0 < f1 or l1 < f0, (n0, n1)        # byte ranges overlap -> live ranges must not
