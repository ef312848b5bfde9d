"""Profiling driver (run under ncu via gpurun): builds the CIFAR-10 DDPM++ engine and runs a few UNet evals."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffpure_b200 import lowering_ncsnpp as L, synthetic
from diffpure_b200.engine import Engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = L.cifar10_cfg()
sd = synthetic.random_state_dict(L.param_shapes(cfg), seed=0)
eng = Engine(L.lower(cfg, sd, B), device=0)
x = torch.rand(B, 3, 32, 32, device="cuda") * 2 - 1
t = torch.full((B,), 50.0, device="cuda")
for _ in range(evals):
    y = eng.unet_forward(x, t)
torch.cuda.synchronize()
if os.environ.get("DP_PRINT_PROFILE"):
    rows = eng.profile_ops(0)
    rows = eng.profile_ops(0)
    prog = eng.program
    gi = 0
    for (kind, ms, fl), op in zip([r for r in rows], [o for o in prog.ops for _ in range(1)]):
        pass
    import collections
    agg = collections.OrderedDict()
    for (kind, ms, fl) in rows:
        agg.setdefault(kind, [0, 0.0, 0.0])
        agg[kind][0] += 1; agg[kind][1] += ms; agg[kind][2] += fl
    for k, v in agg.items():
        print(f"{k:12s} n={v[0]:4d} ms={v[1]:9.3f} TF/s={(v[2]/1e12/(v[1]/1e3) if v[1] else 0):8.1f}")
    # per-gemm detail
    gemm_rows = [(ms, fl) for (kind, ms, fl) in rows if kind == "gemm"]
    gemm_ops = [o for o in prog.ops if o.kind == "gemm"]
    seen = collections.OrderedDict()
    for (ms, fl), o in zip(gemm_rows, gemm_ops):
        a = o.args
        key = (a["B"], a["H"], a["W"], a["N"], tuple((s.C, s.taps) for s in a["a"]), a["batch"], bool(a["resid"]), a["softmax"])
        seen.setdefault(key, [0, 0.0, 0.0])
        seen[key][0] += 1; seen[key][1] += ms; seen[key][2] += fl
    for k, v in sorted(seen.items(), key=lambda kv: -kv[1][1]):
        print(f"gemm B{k[0]} {k[1]}x{k[2]} N={k[3]} segs={k[4]} batch={k[5]} resid={k[6]} sm={k[7]}: n={v[0]} total {v[1]:.3f} ms  {v[2]/1e12/(v[1]/1e3):.1f} TF/s")

    # per-gn_apply detail (bytes = fp32 in + bf16 out (+ raw copies))
    gn_rows = [ms for (kind, ms, fl) in rows if kind == "gn_apply"]
    gn_ops = [o for o in prog.ops if o.kind == "gn_apply"]
    seen = collections.OrderedDict()
    for ms, o in zip(gn_rows, gn_ops):
        a = o.args; C = a["C0"] + a["C1"]; H, W = a["H"], a["W"]
        sc = {0: 1, 1: 4, 2: 0.25}[a["resample"]]
        by = B * H * W * C * (4 + sc * 2 * (2 if a["raw_bf16"] is not None else 1) + (sc * 4 if a["raw_f32"] is not None else 0))
        key = (H, a["C0"], a["C1"], a["resample"], a["raw_bf16"] is not None)
        seen.setdefault(key, [0, 0.0, 0.0]); seen[key][0] += 1; seen[key][1] += ms; seen[key][2] += by
    for k, v in sorted(seen.items(), key=lambda kv: -kv[1][1]):
        print(f"gn H={k[0]} C0={k[1]} C1={k[2]} res={k[3]} raw={k[4]}: n={v[0]} total {v[1]:.3f} ms  {v[2]/1e9/(v[1]/1e3):.0f} GB/s")
