"""Precompute the ORACLE side of the full-size GPU parity cases -> tests/golden/fullsize_oracle.npz (test infrastructure).

The reference-pinned CPU oracles (oracle/adm.py, ddpm_unet.py, ncsnpp.py, sde.py, ddpm_loops.py, ncsnpp_vjp.py; held to the
reference's own modules by tests/golden/*_tiny*.npz and ncsnpp_cifar10_eval.npz) are run here, once, on the seeded
operands of tests/golden_inputs.py with the seeded random-init weights of the real shapes (diffpure_b200/synthetic.py,
oracle/weights.py); the -m gpu tests hold the CUDA path to the stored results instead of recomputing them on the GPU box's
host cores (the full ImageNet network's autograd alone took > 2 minutes there).

    python oracle/make_fullsize_golden.py          # ~3 minutes on 8 cores
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import golden_inputs as GI  # noqa: E402
from diffpure_b200 import lowering_adm as LA, lowering_ddpm as LD, schedule, synthetic  # noqa: E402  (shapes, tables, weights)
from oracle import adm as A, ddpm_loops as OL, ddpm_unet as D, ncsnpp as O, ncsnpp_vjp as V, sde as OS, weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fullsize_oracle.npz")


def main():
    out, t0 = {}, time.time()
    idx = GI.sparse_pixels()
    sp = lambda t: GI.at_pixels(t.detach(), idx).numpy()  # noqa: E731

    sd_adm = synthetic.random_state_dict(LA.param_shapes(LA.imagenet_cfg()), seed=0)
    sd_cel = synthetic.random_state_dict(LD.param_shapes(LD.celeba_cfg()), seed=0)
    sd_cif = weights.make_state_dict(O.param_shapes(O.CIFAR10_CFG), seed=0)
    for name, sd in (("adm", sd_adm), ("celeba", sd_cel), ("cifar", sd_cif)):
        out[f"fp_{name}"] = GI.weights_fingerprint(sd).numpy()
    adm = lambda x, t: A.forward(A.IMAGENET_CFG, sd_adm, x, t)  # noqa: E731
    cel = lambda x, t: D.forward(D.CELEBA_CFG, sd_cel, x, t)  # noqa: E731
    cif = lambda x, t: O.forward(O.CIFAR10_CFG, sd_cif, x, t)  # noqa: E731

    with torch.no_grad():
        # ---- one evaluation at B = 1 (tests/test_gpu_adm_celeba.py::test_full_size_256_eval_vs_oracle)
        x, t = GI.fullsize_eval_inputs()
        out["adm_eval"] = sp(adm(x, t))
        out["celeba_eval"] = sp(cel(x, t.long()))
        print("evals", round(time.time() - t0, 1), "s")
        # ---- 3 steps of the real schedules (::test_full_size_256_three_step_chain_vs_oracle)
        x0, e0, z = GI.fullsize_chain_inputs()
        grid = OS.time_grid(150)
        xx = OS.forward_diffuse(x0, e0, 150)
        for k in range(3):
            tt, h = grid[k], grid[k + 1] - grid[k]
            xx = xx + OS.rev_vpsde_f(adm, "guided_diffusion", tt, xx) * h + \
                OS.rev_vpsde_g(tt, 1)[:, None, None, None] * z[k] * torch.sqrt(h)
        out["chain_adm_vpsde"] = sp(xx)
        _, _, sx, se = schedule.guided_tables(150)
        tab = OL.GuidedTables()
        xx = sx * x0 + se * e0
        for k in range(3):
            xx = OL.guided_p_sample(adm, tab, xx, 149 - k, z[k])
        out["chain_adm_guided"] = sp(xx)
        _, coef, sx, se = schedule.ddpm_tables(100)
        xx = sx * x0 + se * e0
        for k in range(3):
            eps = cel(xx, torch.tensor([99 - k]))
            xx = float(coef[k, 0]) * xx + float(coef[k, 1]) * eps + float(coef[k, 2]) * z[k]
        out["chain_celeba"] = sp(xx)
        print("chains", round(time.time() - t0, 1), "s")
        # ---- CIFAR-10 DDPM++ (tests/test_gpu_parity.py): 30- and 100-step trajectories, one evaluation, input gradient
        x0, e0, z = GI.cifar_traj30_inputs()
        grid = OS.time_grid(100)
        xx = OS.forward_diffuse(x0, e0, 100)
        for k in range(30):
            tt, h = grid[k], grid[k + 1] - grid[k]
            xx = xx + OS.rev_vpsde_f(cif, "score_sde", tt, xx) * h + OS.rev_vpsde_g(tt, 2)[:, None, None, None] * z[k] * torch.sqrt(h)
        out["cifar_traj30"] = xx.numpy()
        x0, e0, z = GI.cifar_traj100_inputs()
        out["cifar_traj100"] = OS.purify_sde(cif, x0[:2], 100, e0[:2], z[:, :2]).numpy()
        x, labels = GI.cifar_pair_eval_inputs()
        out["cifar_pair_eval"] = cif(x[:2], labels[:2]).numpy()
        x, t, go = GI.cifar_vjp_inputs(0)
        out["cifar_vjp"] = V.vjp(O.CIFAR10_CFG, sd_cif, x, t, go).numpy()
        print("cifar", round(time.time() - t0, 1), "s")
    # ---- input gradient of the full ImageNet network (tests/test_gpu_vjp.py::test_adm_fullsize_unet_vjp_vs_autograd)
    x, t, go = GI.fullsize_adm_vjp_inputs()
    x.requires_grad_(True)
    (adm(x, t)[:, :3] * go).sum().backward()
    out["adm_vjp"] = sp(x.grad)
    print("adm vjp", round(time.time() - t0, 1), "s")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()}, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
