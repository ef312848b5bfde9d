"""Shared pieces of the five runners.

* `VPScore`: the continuous VP-SDE quantities (beta(t), the eps -> score conversion for both score-network conventions)
  behind the three "SDE object" classes the reference exposes to torchsde / torchdiffeq (`RevVPSDE`, `VPODE`, `LDSDE`).
* `PurifyRunner`: the control flow every `image_editing_sample` shares around `Engine.purify` -- device choice, the
  `bs_id < 2` image dumps under `log_dir/bs{bs_id}_{tag}`, the `sample_step` passes -- so each runner only states its
  schedule and update kind.
"""
import os
import random

import numpy as np
import torch


def _extract_into_tensor(arr_or_func, timesteps, broadcast_shape):
    """Table lookup or callable evaluated at `timesteps`, broadcast to `broadcast_shape` (runners/diffpure_sde.py:23-39)."""
    if callable(arr_or_func):
        res = arr_or_func(timesteps).float()
    else:
        res = arr_or_func.to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


class VPScore(torch.nn.Module):
    """beta_t = beta_0 + t (beta_1 - beta_0) and score(x, t) from the eps network, on flattened states (B, C*H*W)."""

    def __init__(self, model, score_type, beta_min, beta_max, N, img_shape, model_kwargs):
        super().__init__()
        self.model = model
        self.score_type = score_type
        self.model_kwargs = model_kwargs
        self.img_shape = img_shape
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1. - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1. - self.alphas_cumprod)
        self.alphas_cumprod_cont = lambda t: torch.exp(-0.5 * (beta_max - beta_min) * t ** 2 - beta_min * t)
        self.sqrt_1m_alphas_cumprod_neg_recip_cont = lambda t: -1. / torch.sqrt(1. - self.alphas_cumprod_cont(t))

    def _scale_timesteps(self, t):
        assert torch.all(t <= 1) and torch.all(t >= 0), f't has to be in [0, 1], but get {t} with shape {t.shape}'
        return (t.float() * self.N).long()

    def beta(self, t):
        return self.beta_0 + t * (self.beta_1 - self.beta_0)

    def score(self, t, x):
        """t: (B,) forward time, x: (B, D) -> (B, D)."""
        assert x.ndim == 2 and np.prod(self.img_shape) == x.shape[1], x.shape
        x_img = x.view(-1, *self.img_shape)
        if self.score_type == 'guided_diffusion':      # discrete-time eps(+var) network, eps -> score with sigma(t)
            eps = self.model(x_img, self._scale_timesteps(t))
            eps, _ = torch.split(eps, self.img_shape[0], dim=1)
            scale = _extract_into_tensor(self.sqrt_1m_alphas_cumprod_neg_recip_cont, t, x.shape)
            return scale * eps.reshape(x.shape[0], -1)
        if self.score_type == 'score_sde':             # continuous-time network fed t*999 (score_sde/models/utils.py:149)
            eps = self.model(x_img, t * 999)
            log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
            std = torch.sqrt(1. - torch.exp(2. * log_mean_coeff))                      # sde_lib.py:149-153
            return (-eps / std[:, None, None, None]).reshape(x.shape[0], -1)
        raise NotImplementedError(f'Unknown score type in RevVPSDE: {self.score_type}!')


class PurifyWithGrad(torch.autograd.Function):
    """x0 (and the Langevin anchor) -> purified x for the linear updates
        x_{k+1} = c0_k x_k + c1_k eps(x_k, cond_k) + c2_k z_k (+ c3_k anchor),    x_0 = sx x0 + se e,
    differentiable in x0 and the anchor. The reference differentiates with torchsde's / torchdiffeq's continuous adjoint
    (runners/diffpure_sde.py:233-239, diffpure_ode.py:230-238, diffpure_ldsde.py:240-243: an augmented system solved
    backwards in time); here the backward pass is the exact gradient of the discrete loop the engine runs
    (discretise-then-differentiate): the forward pass records the states x_k, the backward pass replays them through the
    engine's UNet vector-Jacobian program (`dp_unet_vjp`):
        lambda_k = c0_k lambda_{k+1} + J_k^T (c1_k lambda_{k+1}),   dL/d anchor = sum_k c3_k lambda_{k+1},   dL/dx0 = sx lambda_0.
    Restated on the CPU by oracle/ncsnpp_vjp.py:purify_sde_vjp (held to torch.autograd through the oracle loop)."""

    @staticmethod
    def forward(ctx, x0, anchor, model, cond, coef, sx, se, e, step_noise, seed, sample_offset, update_kind):
        dev = x0.device
        eng = model.engine_for(x0.shape[0], dev)
        states = torch.empty((len(cond) + 1,) + tuple(x0.shape), device=dev, dtype=torch.float32)
        kw = dict(anchor=anchor.detach()) if anchor is not None else {}
        out = eng.purify(x0.detach(), cond, coef, sx, se, update_kind=update_kind, init_noise=e, step_noise=step_noise,
                         seed=seed, sample_offset=sample_offset, states=states, **kw)
        ctx.model, ctx.cond, ctx.coef, ctx.sx, ctx.states = model, cond, np.asarray(coef, dtype=np.float32), sx, states
        ctx.has_anchor = anchor is not None
        return out

    @staticmethod
    def backward(ctx, g):
        states, cond, coef = ctx.states, ctx.cond, ctx.coef
        B, dev = states.shape[1], states.device
        veng = ctx.model.engine_for(B, dev, vjp=True)
        lam = g.contiguous().float()
        ga = torch.zeros_like(lam) if ctx.has_anchor else None
        for k in reversed(range(len(cond))):
            if ga is not None:
                ga = ga + float(coef[k, 3]) * lam
            ck = torch.full((B,), float(cond[k]), device=dev)
            lam = float(coef[k, 0]) * lam + veng.unet_vjp(states[k], ck, float(coef[k, 1]) * lam)
        return (ctx.sx * lam, ga) + (None,) * 10


class _Dump:
    """The reference's per-batch image dumps (only for the first two batches)."""

    def __init__(self, args, bs_id, tag):
        if tag is None:
            tag = 'rnd' + str(random.randint(0, 10000))
        self.dir = os.path.join(args.log_dir, 'bs' + str(bs_id) + '_' + tag)
        self.on = bs_id < 2 and getattr(args, "save_images", True)
        if self.on:
            os.makedirs(self.dir, exist_ok=True)

    def image(self, name, x):
        if self.on:
            import torchvision.utils as tvu
            tvu.save_image((x + 1) * 0.5, os.path.join(self.dir, name))

    def tensor(self, name, x):
        if self.on:
            torch.save(x, os.path.join(self.dir, name))


class PurifyRunner(torch.nn.Module):
    """Base of the runner classes: `self.model` is a `ScoreModel` (engine factory)."""

    differentiable_error = None     # message raised for inputs that require grad (None: gradients flow / no_grad runner)
    device_from_input = False       # Diffusion takes the device from the input image (runners/diffpure_ddpm.py:126,129)

    def _setup(self, args, config, device):
        self.args = args
        self.config = config
        if device is None:
            device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        self.sample_offset = 0      # global index of this shard's first sample (multi-GPU sharding)
        self.last_seed = None

    def _call_seed(self, seed, it):
        """Seed of one pass: drawn from NumPy's global RNG (as torchsde's BrownianInterval does without entropy) or
        derived from the caller's."""
        s = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed) + it
        self.last_seed = s
        return s

    def _open(self, img, bs_id, tag):
        assert isinstance(img, torch.Tensor)
        if self.differentiable_error and torch.is_grad_enabled() and img.requires_grad:
            raise NotImplementedError(self.differentiable_error)
        assert img.ndim == 4, img.ndim
        if self.device_from_input:
            dev = img.device if img.device.type == "cuda" else self.device
        elif self.device.type == "cuda" and self.device.index is None and img.device.type == "cuda":
            dev = img.device        # a generic 'cuda' device = the caller's current one: nn.DataParallel replicas receive
                                    # their chunk on their own GPU (eval_sde_adv.py:227-229) and must run there
        else:
            dev = self.device if self.device.type == "cuda" else img.device
        x0 = img.to(dev)
        dump = _Dump(self.args, bs_id, tag)
        dump.image('original_input.png', x0)
        return x0, dev, dump

    _fuse_kw = {}                   # pre / post steps fused into this call's dp_purify (see purify_unit_range)

    def purify_unit_range(self, x01, out_hw=None, out_norm=None, bs_id=2, tag=None, **kw):
        """The caller's pre / post steps fused into the engine call (SDE_Adv_Model.forward, eval_sde_adv.py:73-89):
        x01 in [0,1] at any spatial size -> bilinear resize to the model grid, (x - 0.5) * 2, forward diffusion, the loop,
        bilinear resize to `out_hw`, (x + 1) / 2 and optionally the classifier normalisation `out_norm = (mean, std)`
        (utils.py:144-153) -- no eager kernels either side of the loop. Needs sample_step == 1 and no image dumps
        (bs_id >= 2 or save_images off); forward only."""
        if self.args.sample_step != 1:
            raise ValueError("purify_unit_range needs sample_step == 1")
        if torch.is_grad_enabled() and x01.requires_grad:
            raise ValueError("purify_unit_range is forward-only; use image_editing_sample for gradients")
        if bs_id < 2 and getattr(self.args, "save_images", True):
            raise ValueError("purify_unit_range does not write the bs_id < 2 image dumps")
        self._fuse_kw = dict(in_unit_range=True, out_hw=tuple(out_hw) if out_hw else None, out_unit_range=True,
                             out_norm=out_norm)
        try:
            return self.image_editing_sample(x01, bs_id=bs_id, tag=tag, **kw)
        finally:
            self._fuse_kw = {}

    def _init_noise(self, x, init_noise, dev):
        """The forward-diffusion draw (reference: torch.randn_like(x)); fused calls leave it to the engine's generator."""
        if init_noise is not None:
            return init_noise.to(dev)
        return None if self._fuse_kw else torch.randn_like(x)

    def _wants_grad(self, x):
        """True when the caller differentiates through the loop (white-box attacks); raises for networks whose
        input-gradient program does not exist yet instead of silently detaching."""
        if not (torch.is_grad_enabled() and x.requires_grad):
            return False
        if getattr(self.model, "_lower_vjp", None) is None and not hasattr(self.model, "vjp_ok"):
            raise NotImplementedError(
                "diffpure_b200: backward through the purification loop is implemented for the DDPM++ (CIFAR-10) and ADM "
                "(ImageNet) networks; wrap the call in torch.no_grad() / detach the input for this network")
        return True

    def _passes(self, x0, dump, one_pass):
        """`sample_step` purification passes, each starting from the previous one's output; returns their concatenation."""
        xs = []
        for it in range(self.args.sample_step):
            x0 = one_pass(it, x0)
            dump.tensor(f'samples_{it}.pth', x0)
            dump.image(f'samples_{it}.png', x0)
            xs.append(x0)
        return torch.cat(xs, dim=0)
