"""Seeded operands of the fixtures that do not store them (the generating script draws them the same way)."""
import torch

from oracle import sde as OS


def adm_vpsde_inputs(input_seed, t_star, B=2, S=64):
    """x0, e0, z of tests/golden/adm_tiny_vpsde.npz (oracle/make_golden.py:golden_adm_vpsde)."""
    g = torch.Generator().manual_seed(int(input_seed))
    x0 = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    e0 = torch.randn(B, 3, S, S, generator=g)
    z = torch.randn(OS.num_steps(int(t_star)), B, 3, S, S, generator=g)
    return x0, e0, z


def respaced_chain_inputs(input_seed, B=2, S=64):
    """x0, e0 of the respaced p_sample chain in tests/golden/guided_schedules.npz (make_golden.py:golden_guided_schedules)."""
    g = torch.Generator().manual_seed(int(input_seed))
    x0 = torch.rand(B, 3, S, S, generator=g) * 2 - 1
    e0 = torch.randn(B, 3, S, S, generator=g)
    return x0, e0


ADM_TINY_REF_CONFIG = dict(image_size=64, num_channels=64, num_res_blocks=1, attention_resolutions="32,16,8",
                           num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True, learn_sigma=True,
                           class_cond=False, diffusion_steps=1000, channel_mult="", timestep_respacing="1000",
                           noise_schedule="linear", rescale_timesteps=True)


# ---------------------------------------------------------------------------------------------------------------------
# Full-size cases whose oracle side is precomputed (oracle/make_fullsize_golden.py -> tests/golden/fullsize_oracle.npz):
# the oracle needs seconds to minutes per case on a many-core host, the GPU side milliseconds. Both sides draw the operands
# from these seeded builders; 256x256 results are stored at a seeded random eighth of the pixel positions (every 128-pixel
# GEMM tile is hit ~16 times), CIFAR-sized results in full.
# ---------------------------------------------------------------------------------------------------------------------
def sparse_pixels(hw=256 * 256, keep_one_in=8, seed=4242):
    g = torch.Generator().manual_seed(seed)
    return torch.randperm(hw, generator=g)[:hw // keep_one_in].sort().values


def at_pixels(t, idx):
    """[B, C, H, W] -> [B, C, len(idx)]"""
    return t.reshape(t.shape[0], t.shape[1], -1)[:, :, idx]


def fullsize_eval_inputs():
    g = torch.Generator().manual_seed(0)
    return torch.rand(1, 3, 256, 256, generator=g) * 2 - 1, torch.tensor([77.0])


def fullsize_chain_inputs():
    g = torch.Generator().manual_seed(21)
    x0 = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    e0 = torch.randn(1, 3, 256, 256, generator=g)
    z = torch.randn(3, 1, 3, 256, 256, generator=g)
    return x0, e0, z


def fullsize_adm_vjp_inputs():
    g = torch.Generator().manual_seed(3)
    x = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    go = torch.randn(1, 3, 256, 256, generator=g)
    return x, torch.tensor([77.0]), go


def cifar_traj30_inputs():
    g = torch.Generator().manual_seed(5)
    B, steps = 2, 30
    return (torch.rand(B, 3, 32, 32, generator=g) * 2 - 1, torch.randn(B, 3, 32, 32, generator=g),
            torch.randn(steps, B, 3, 32, 32, generator=g))


def cifar_traj100_inputs():
    g = torch.Generator().manual_seed(6)
    B, steps = 96, 100
    return (torch.rand(B, 3, 32, 32, generator=g) * 2 - 1, torch.randn(B, 3, 32, 32, generator=g),
            torch.randn(steps, B, 3, 32, 32, generator=g))


def cifar_pair_eval_inputs():
    g = torch.Generator().manual_seed(15)
    return torch.rand(96, 3, 32, 32, generator=g) * 2 - 1, torch.full((96,), 37.0)


def cifar_vjp_inputs(seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(2, 3, 32, 32, generator=g) * 2 - 1
    return x, torch.tensor([37.0, 512.0]), torch.randn(2, 3, 32, 32, generator=g)


def weights_fingerprint(sd, n=6):
    """A few float64 sums over the seeded random-init weights: the precomputed oracle results only hold for these values."""
    keys = sorted(sd.keys())
    pick = [keys[(i * len(keys)) // n] for i in range(n)]
    return torch.tensor([sd[k].double().abs().sum().item() for k in pick], dtype=torch.float64)
