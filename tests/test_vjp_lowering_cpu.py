"""CPU: the forward + data-gradient program (`lowering_ncsnpp.lower_vjp`, the program of dp_unet_vjp) replayed by the
program interpreter reproduces oracle/ncsnpp_vjp.py (itself held to torch.autograd) to fp32 round-off; with bf16 operand
emulation it stays inside the tolerance stated for the tensor-core gradient path (3e-2: the bf16 rounding of the frozen
weights alone moves the input gradient by 1.4e-2 on these networks)."""
import pytest
import torch

from diffpure_b200 import lowering_ncsnpp as L
from oracle import ncsnpp as O, ncsnpp_vjp as V, weights
from program_interp import Interp

CASES = [
    ("small-attn", O.tiny_cfg(64, (1, 2), 1, (8,), 16), 2, 1),            # T = 64: attn_small / attn_small_bwd
    ("tc-attn-updown", O.tiny_cfg(64, (1, 2, 2), 2, (16,), 32), 2, 2),    # T = 256: GEMM attention + softmax_bwd + transposes
]


@pytest.mark.parametrize("name,cfg,B,seed", CASES)
def test_vjp_program_matches_the_vjp_oracle(name, cfg, B, seed):
    sd = weights.make_state_dict(O.param_shapes(cfg), seed=seed)
    g = torch.Generator().manual_seed(seed)
    S = cfg.image_size
    x = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    t = torch.tensor([37.0, 512.0])
    go = torch.randn(B, 3, S, S, generator=g)
    with torch.no_grad():
        ref = V.vjp(cfg, sd, x, t, go)
    prog = L.lower_vjp(cfg, sd, B)
    kinds = {o.kind for o in prog.ops}
    assert {"grad_in", "gn_bwd", "gemm", "update"} <= kinds
    assert "attn_small_bwd" in kinds                                 # the middle block's short sequence in both cases
    assert ({"softmax_bwd", "transpose"} <= kinds) == (name != "small-attn")
    got = Interp(prog, emulate_bf16=False).run_vjp(x, t, go)
    assert ((got - ref).norm() / ref.norm()).item() < 1e-4
    got16 = Interp(prog, emulate_bf16=True).run_vjp(x, t, go)
    assert ((got16 - ref).norm() / ref.norm()).item() < 3e-2


def test_dgrad_weight_packing():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(24, 16, 3, 3, generator=g)
    gy = torch.randn(2, 24, 8, 8, generator=g)
    ref = V.conv_dgrad(gy, w)                                       # [2, 16, 8, 8]
    wp = L.pack_dgrad3x3(w)                                         # [16, 9*24], K = (ky*3+kx)*24 + co
    cols = torch.nn.functional.unfold(gy, 3, padding=1)            # [2, 24*9, 64] with K = co*9 + tap
    cols = cols.reshape(2, 24, 9, 64).permute(0, 2, 1, 3).reshape(2, 9 * 24, 64)
    got = torch.einsum("nk,bkp->bnp", wp, cols).reshape(2, 16, 8, 8)
    assert torch.allclose(got, ref, atol=1e-4)


def test_adm_vjp_program_matches_autograd_on_the_oracle():
    """guided_diffusion UNet (scale-shift norm, resblock up / down, multi-head attention at T = 1024 / 256 / 64): the program of
    `lowering_adm.lower_vjp` against torch.autograd through oracle/adm.py (pinned to the reference modules), for the
    gradient wrt the eps half of the output (what the VP-SDE path of runners/diffpure_sde.py:96-122 reads)."""
    from diffpure_b200 import lowering_adm as LA
    from oracle import adm as A
    cfg = A.tiny_cfg(64, 64, (1, 2, 3, 4), 1, (32, 16, 8))
    B = 2
    sd = weights.make_state_dict(A.param_shapes(cfg), seed=5)
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1).requires_grad_(True)
    t = torch.tensor([37.0, 512.0])
    go = torch.randn(B, 3, 64, 64, generator=g)
    (A.forward(cfg, sd, x, t)[:, :3] * go).sum().backward()
    ref = x.grad
    prog = LA.lower_vjp(cfg, sd, B)
    kinds = [o.kind for o in prog.ops]
    assert {"grad_in", "gn_bwd", "gemm", "update", "attn_small_bwd", "softmax_bwd", "transpose"} <= set(kinds)
    assert any(o.kind == "gn_bwd" and o.args["film"] is not None for o in prog.ops)            # scale-shift norm
    assert any(o.kind == "softmax_bwd" and o.args["rowsum"] is None for o in prog.ops)         # T = 1024: normalised P
    got = Interp(prog, emulate_bf16=False).run_vjp(x.detach(), t, go)
    assert ((got - ref).norm() / ref.norm()).item() < 1e-4
    got16 = Interp(prog, emulate_bf16=True).run_vjp(x.detach(), t, go)
    assert ((got16 - ref).norm() / ref.norm()).item() < 3e-2
