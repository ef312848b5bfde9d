"""diffpure_b200 -- B200-native (sm_100a) engine for DiffPure's reverse-SDE purification hot path.

Package layout (only what the path needs):
  csrc/                hand-written CUDA (tcgen05/TMA implicit-GEMM, fused GroupNorm/SiLU, fused SDE update)
                       and the C-ABI engine -> libdiffpure_b200.so  (include/diffpure_b200.h)
  lib.py               ctypes binding; raises if the library is missing (no CPU fallback)
  program.py           the op program the UNets are lowered to
  lowering_*.py        DDPM++ / ADM / DDPM UNet -> program (reference parameter names)
  engine.py            materialises a program through the C ABI, runs UNet evals and the device-side loop
  schedule.py          per-step scalar tables in the reference's fp32 op order
  runners/             RevVPSDE, RevGuidedDiffusion, GuidedDiffusion, Diffusion with the reference's API
"""
__version__ = "0.1.0"
