"""GPU check + timing of the full-size 256x256 models (run under gpurun): one UNet eval vs the CPU oracle at B=1,
then per-op profile and a short purification at the BASELINE batch."""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from diffpure_b200 import lib, schedule, synthetic
from diffpure_b200.engine import Engine

which, B = sys.argv[1], int(sys.argv[2])
check = len(sys.argv) > 3 and sys.argv[3] == "check"
torch.set_grad_enabled(False)
if which == "adm":
    from diffpure_b200 import lowering_adm as L
    cfg = L.imagenet_cfg(); F = 2239.67e9; steps = 150
    cond, coef, sx, se = schedule.guided_tables(150); kind = lib.DP_UPDATE_LEARNED_RANGE
else:
    from diffpure_b200 import lowering_ddpm as L
    cfg = L.celeba_cfg(); F = 497.03e9; steps = 100
    cond, coef, sx, se = schedule.ddpm_tables(100); kind = lib.DP_UPDATE_LINEAR
sd = synthetic.random_state_dict(L.param_shapes(cfg), seed=0)
if check:
    if which == "adm":
        from oracle import adm as O
        ocfg = O.IMAGENET_CFG
    else:
        from oracle import ddpm_unet as O
        ocfg = O.CELEBA_CFG
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    t = torch.tensor([77.0])
    t0 = time.time(); y = O.forward(ocfg, sd, x, t if which == "adm" else t.long()); print("oracle eval %.1fs" % (time.time() - t0))
    eng = Engine(L.lower(cfg, sd, 1), device=0)
    yg = eng.unet_forward(x.cuda(), t.cuda()).cpu()
    print(which, "full-size eval rel-L2 %.3e (|y| %.3f)" % (((yg - y).norm() / y.norm()).item(), y.abs().mean().item()))
    eng.close()
t0 = time.time()
eng = Engine(L.lower(cfg, sd, B), device=0)
print("engine B=%d built in %.1fs, %d launches/eval, act %.1f GB, weights %.2f GB" % (B, time.time() - t0, eng.launches_per_eval, eng.act_bytes / 1e9, eng.const_bytes / 1e9))
x = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1
for _ in range(2):
    rows = eng.profile_ops(1)
agg = collections.OrderedDict()
for k, ms, fl in rows:
    agg.setdefault(k, [0, 0.0, 0.0]); agg[k][0] += 1; agg[k][1] += ms; agg[k][2] += fl
tot = sum(v[1] for v in agg.values())
for k, v in agg.items():
    print(f"{k:12s} n={v[0]:4d} ms={v[1]:9.3f} TF/s={(v[2] / 1e12 / (v[1] / 1e3) if v[1] else 0):8.1f}")
print("eval total %.2f ms -> %.1f TF/s algorithmic (%.3f of 1416.5)" % (tot, B * F / tot / 1e9, B * F / tot / 1e9 / 1416.5))
gemm_rows = [(ms, fl) for (kind, ms, fl) in rows if kind == "gemm"]
gemm_ops = [o for o in eng.program.ops if o.kind == "gemm"]
seen = collections.OrderedDict()
for (ms, fl), o in zip(gemm_rows, gemm_ops):
    a = o.args
    key = (a["B"], a["H"], a["W"], a["N"], tuple((s.C, s.taps) for s in a["a"]), a["batch"], bool(a["resid"]), a["softmax"], a["out_bf16"] is not None)
    seen.setdefault(key, [0, 0.0, 0.0]); seen[key][0] += 1; seen[key][1] += ms; seen[key][2] += fl
for k, v in sorted(seen.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"gemm B{k[0]} {k[1]}x{k[2]} N={k[3]} segs={k[4]} batch={k[5]} resid={k[6]} sm={k[7]} bf16={k[8]}: n={v[0]} total {v[1]:.3f} ms  {v[2]/1e12/(v[1]/1e3):.1f} TF/s")
n = 6
torch.cuda.synchronize(); t0 = time.time()
out = eng.purify(x, cond[:n], coef[:n], sx, se, update_kind=kind, seed=1)
torch.cuda.synchronize(); dt = time.time() - t0
print("%d steps in %.3fs -> %.1f ms/step -> full %d-step purification %.2f s -> %.3f img/s" % (n, dt, dt / n * 1e3, steps, dt / n * steps, B / (dt / n * steps)))
assert torch.isfinite(out).all()
