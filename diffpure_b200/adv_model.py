"""SDE_Adv_Model with the reference's interface (eval_sde_adv.py:34-93; the BPDA script's variant, eval_sde_adv_bpda.py:53-117,
as SDE_Adv_Model_BPDA) on the B200 engine.

The reference class works unmodified on `diffpure_b200.runners` (tests/test_siblings_cpu.py); this one additionally fuses
its eager pre / post steps into the engine call when no gradient is requested:

    reference forward                                        here (no-grad / BPDA / clean + robust accuracy evaluation)
    F.interpolate(x, 256, bilinear)        [ImageNet]   \\
    (x - 0.5) * 2                                        |   ONE dp_purify: the resize, both range maps, the forward
    runner.image_editing_sample(...)                     |   diffusion and the classifier's (x - mu) / sigma are kernels
    F.interpolate(x_re, 224, bilinear)     [ImageNet]    |   either side of the device-resident loop
    classifier((x_re + 1) * 0.5)  -> (x - mu) / sigma   /    classifier.resnet(...) on the normalised tensor

With a requires_grad input (white-box attacks) the torch ops of the reference run around the differentiable runner call.
"""
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


def _runner_for(args, config, state_dict=None):
    kw = {} if state_dict is None else {"state_dict": state_dict}
    if args.diffusion_type == 'ddpm':
        from .runners.diffpure_guided import GuidedDiffusion
        return GuidedDiffusion(args, config, device=config.device, **kw)
    if args.diffusion_type == 'sde':
        from .runners.diffpure_sde import RevGuidedDiffusion
        return RevGuidedDiffusion(args, config, device=config.device, **kw)
    if args.diffusion_type == 'ode':
        from .runners.diffpure_ode import OdeGuidedDiffusion
        return OdeGuidedDiffusion(args, config, device=config.device, **kw)
    if args.diffusion_type == 'ldsde':
        from .runners.diffpure_ldsde import LDGuidedDiffusion
        return LDGuidedDiffusion(args, config, device=config.device, **kw)
    if args.diffusion_type == 'celebahq-ddpm':
        from .runners.diffpure_ddpm import Diffusion
        return Diffusion(args, config, device=config.device, **kw)
    raise NotImplementedError('unknown diffusion type')


class SDE_Adv_Model(nn.Module):
    def __init__(self, args, config, classifier=None, state_dict=None):
        """Reference arguments (L35) plus: `classifier` (the reference builds it with utils.get_image_classifier from
        pretrained files that do not exist offline) and `state_dict` for the score network."""
        super().__init__()
        self.args = args
        self._device = config.device
        if classifier is None:
            from utils import get_image_classifier          # the reference's factory, when its tree is importable
            classifier = get_image_classifier(args.classifier_name)
        self.classifier = classifier.to(config.device)
        print(f'diffusion_type: {args.diffusion_type}')
        self.runner = _runner_for(args, config, state_dict)
        self.register_buffer('counter', torch.zeros(1, device=config.device))
        self._count = 0     # host copy of `counter` (the reference reads the device buffer with .item() every call)
        self.tag = None

    def reset_counter(self):
        self.counter = torch.zeros(1, dtype=torch.int, device=self._device)
        self._count = 0

    def set_tag(self, tag=None):
        self.tag = tag

    def _fusable(self, x):
        return (not (torch.is_grad_enabled() and x.requires_grad)) and self.args.sample_step == 1 and \
            x.is_cuda and hasattr(self.runner, "purify_unit_range") and self.args.diffusion_type != 'ldsde' and \
            (self._count >= 2 or not getattr(self.args, "save_images", True))

    def _purify(self, x, norm=None):
        """The purification half of the reference's forward (eval_sde_adv.py:68-89, eval_sde_adv_bpda.py:84-104): returns
        (images in [0,1] at the input's size, fused). On the fused path `norm = (mean, std)` is applied as well."""
        counter = self._count
        if counter % 5 == 0:
            print(f'diffusion times: {counter}')
        imagenet = 'imagenet' in self.args.domain
        fused = self._fusable(x)
        start_time = time.time()
        if fused:
            x01 = self.runner.purify_unit_range(x, out_hw=tuple(x.shape[2:]) if imagenet else None, out_norm=norm,
                                                bs_id=max(counter, 2), tag=self.tag)
            shape_in = (x.shape[0], 3, 256, 256) if imagenet else tuple(x.shape)
        else:
            if imagenet:    # imagenet [3, 224, 224] -> [3, 256, 256] -> [3, 224, 224]
                x = F.interpolate(x, size=(256, 256), mode='bilinear', align_corners=False)
            shape_in = tuple(x.shape)
            x_re = self.runner.image_editing_sample((x - 0.5) * 2, bs_id=counter, tag=self.tag)
            if imagenet:
                x_re = F.interpolate(x_re, size=(224, 224), mode='bilinear', align_corners=False)
            x01 = (x_re + 1) * 0.5
        minutes, seconds = divmod(time.time() - start_time, 60)
        if counter % 5 == 0:
            print(f'x shape (before diffusion models): {torch.Size(shape_in)}')
            print(f'x shape (before classifier): {x01.shape}')
            print("Sampling time per batch: {:0>2}:{:05.2f}".format(int(minutes), seconds))
        self.counter += 1
        self._count += 1
        return x01, fused

    def forward(self, x):
        # classifier wrappers of the reference (utils.py:144-153) normalise inside forward: on the fused path they get the
        # already normalised tensor when they expose (mu, sigma, resnet); any other classifier gets the [0,1] images
        norm, head = None, self.classifier
        if all(hasattr(self.classifier, a) for a in ("mu", "sigma", "resnet")):
            norm = (self.classifier.mu.flatten().tolist(), self.classifier.sigma.flatten().tolist())
            head = self.classifier.resnet
        x01, fused = self._purify(x, norm)
        return head(x01) if fused else self.classifier(x01)


class SDE_Adv_Model_BPDA(SDE_Adv_Model):
    """The BPDA + EOT script's variant of the class (eval_sde_adv_bpda.py:53-117): the classifier is called `resnet`, and
    `forward(x, mode)` selects 'purify' (images in [0,1]), 'classify' (x in [0,1]) or 'purify_and_classify'. Only
    'ddpm', 'sde' and 'celebahq-ddpm' runners exist there (L63-70). BPDA never differentiates through `purify`
    (bpda_eot/bpda_eot_attack.py: the purified batch is detached and the classifier's gradient is taken at it), so every
    purification takes the fused path: resize, range maps and the loop in ONE dp_purify call."""

    def __init__(self, args, config, classifier=None, state_dict=None):
        if args.diffusion_type not in ('ddpm', 'sde', 'celebahq-ddpm'):
            raise NotImplementedError('unknown diffusion type')
        super().__init__(args, config, classifier=classifier, state_dict=state_dict)

    @property
    def resnet(self):
        return self.classifier

    def purify(self, x):
        return self._purify(x, None)[0]

    def forward(self, x, mode='purify_and_classify'):
        if mode == 'purify':
            return self.purify(x)
        if mode == 'classify':
            return self.classifier(x)  # x in [0, 1]
        if mode == 'purify_and_classify':
            return self.classifier(self.purify(x))  # the wrapper normalises the [0, 1] images itself
        raise NotImplementedError(f'unknown mode: {mode}')
