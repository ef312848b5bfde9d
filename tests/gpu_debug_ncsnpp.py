"""GPU debugging aid (run under gpurun): runs a tiny DDPM++ through the engine with pooling disabled and
compares every intermediate tensor against the CPU interpreter of the same program."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ncsnpp as O, weights  # noqa: E402
from diffpure_b200 import lowering_ncsnpp as L  # noqa: E402
from diffpure_b200.engine import Engine  # noqa: E402
from program_interp import Interp  # noqa: E402


def written_tensors(op):
    keys = {"embed": ["out"], "gemm": ["out_f32", "out_bf16", "stats", "rowsum_out"],
            "gn_apply": ["out_bf16", "raw_bf16", "raw_f32"], "conv_in": ["out", "stats"],
            "attn_small": ["out"]}[op.kind]
    return [op.args[k] for k in keys if op.args.get(k) is not None]


def main():
    torch.manual_seed(0)
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    if which == "small":
        cfg, B = O.tiny_cfg(nf=64, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(8,), image_size=16), 3
    elif which == "tc":
        cfg, B = O.tiny_cfg(nf=64, ch_mult=(1, 2, 2), num_res_blocks=1, attn_resolutions=(16,), image_size=32), 2
    else:
        cfg, B = O.CIFAR10_CFG, 2
    sd = weights.make_state_dict(O.param_shapes(cfg), seed=1)
    x = torch.rand(B, 3, cfg.image_size, cfg.image_size) * 2 - 1
    t = torch.rand(B) * 999
    y = O.forward(cfg, sd, x, t)
    prog = L.lower(cfg, sd, B)
    it = Interp(prog, emulate_bf16=True)
    yi = it.run(x, t)
    t0 = time.time()
    eng = Engine(prog, device=0, pool=False)
    print(f"engine built in {time.time() - t0:.1f}s, {len(prog.ops)} ops, {eng.launches_per_eval} launches/eval")
    yg = eng.unet_forward(x.cuda(), t.cuda()).cpu()
    print("engine vs oracle  rel-L2 %.3e  max %.3e" % (((yg - y).norm() / y.norm()).item(), (yg - y).abs().max().item()))
    print("interp vs oracle  rel-L2 %.3e" % ((yi - y).norm() / y.norm()).item())
    print("engine vs interp  rel-L2 %.3e" % ((yg - yi).norm() / yi.norm()).item())
    bad = 0
    for i, op in enumerate(prog.ops):
        for v in written_tensors(op):
            tname = v.tensor.name
            g = eng.read_tensor(tname)
            c = it.mem[v.tensor.index]
            n = min(g.numel(), c.numel())
            if tname.endswith(".stats"):
                # rows beyond the valid segments are never read
                pass
            den = c[:n].norm().item() + 1e-12
            rel = (g[:n] - c[:n]).norm().item() / den
            flag = "" if rel < 3e-2 else "   <-- MISMATCH"
            if flag or os.environ.get("DP_VERBOSE"):
                print(f"op {i:3d} {op.kind:10s} {tname:28s} rel {rel:.3e} |ref| {den:.3e}{flag}")
            if flag:
                bad += 1
                if bad > 8:
                    return 1
    print("intermediate check done, mismatches:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
