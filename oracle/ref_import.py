"""Import the *reference's own* Python modules from /root/reference -- build-container only.

Used to pin the oracle restatements (oracle/check_against_reference.py) and to generate the committed
golden vectors (oracle/make_golden.py). The GPU box has no /root/reference: nothing that runs there
may import this module.

Third-party packages the reference imports but that are absent offline are replaced by inert stubs
(SURVEY.md section 8c): torchsde, torchdiffeq, autoattack, robustbench, lmdb, matplotlib.
"""
import os
import sys
import types
from types import SimpleNamespace

REF_ROOT = os.environ.get("DIFFPURE_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "runners"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(euler_shim=None):
    """Put the reference on sys.path with stubs for its missing third-party imports."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ts = _stub("torchsde")
    if euler_shim is not None:
        ts.sdeint_adjoint = euler_shim
        ts.sdeint = euler_shim
    if not hasattr(ts, "BrownianInterval"):
        ts.BrownianInterval = lambda **kw: None
    _stub("torchdiffeq", odeint_adjoint=None, odeint=None)
    _stub("autoattack", AutoAttack=object)
    rb = _stub("robustbench")
    rb.load_model = lambda *a, **k: None
    _stub("lmdb")
    mpl = _stub("matplotlib")
    mpl.use = lambda *a, **k: None
    _stub("matplotlib.pyplot")
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/torch_extensions_ref")


def to_namespace(d):
    if isinstance(d, dict):
        return SimpleNamespace(**{k: to_namespace(v) for k, v in d.items()})
    return d


def load_config(name):
    import yaml
    with open(os.path.join(REF_ROOT, "configs", name)) as f:
        return to_namespace(yaml.safe_load(f))


def build_ncsnpp(overrides=None):
    """Reference NCSNpp (score_sde/models/ncsnpp.py) on CPU. First import JIT-builds score_sde/op (~2 min)."""
    install()
    from score_sde.models import ncsnpp  # noqa: F401  (registers the model)
    from score_sde.models import utils as mutils
    cfg = load_config("cifar10.yml")
    if overrides:
        for k, v in overrides.items():
            tgt, key = (cfg.data, k[5:]) if k.startswith("data.") else (cfg.model, k)
            setattr(tgt, key, v)
    import torch
    cfg.device = torch.device("cpu")
    model = mutils.create_model(cfg)
    return model.eval(), cfg


def build_adm(num_channels=256, image_size=256, num_res_blocks=2, attention_resolutions="32,16,8", use_fp16=False):
    """Reference ADM UNet (guided_diffusion/unet.py via script_util.create_model_and_diffusion) on CPU."""
    install()
    from guided_diffusion.script_util import create_model_and_diffusion, model_and_diffusion_defaults
    cfg = load_config("imagenet.yml")
    mc = model_and_diffusion_defaults()
    mc.update(vars(cfg.model))
    mc.update(num_channels=num_channels, image_size=image_size, num_res_blocks=num_res_blocks,
              attention_resolutions=attention_resolutions, use_fp16=use_fp16)
    model, diffusion = create_model_and_diffusion(**mc)
    if use_fp16:
        model.convert_to_fp16()
    return model.eval(), diffusion, mc


def build_celeba(overrides=None):
    """Reference CelebA-HQ DDPM UNet (ddpm/unet_ddpm.py) on CPU."""
    install()
    from ddpm.unet_ddpm import Model
    cfg = load_config("celeba.yml")
    if overrides:
        for k, v in overrides.items():
            tgt, key = (cfg.data, k[5:]) if k.startswith("data.") else (cfg.model, k)
            setattr(tgt, key, v)
    return Model(cfg).eval(), cfg
