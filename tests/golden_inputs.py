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
