"""Engine-backed score models: drop-in for the reference's UNet modules (`model(x, t) -> eps`).

The object is an nn.Module without parameters so that `.to()`, `.eval()` and nn.DataParallel wrapping of
the runners keep working (eval_sde_adv.py:227-229); the weights live in the engine's device blob.
"""
import torch

from .engine import Engine


class ScoreModel(torch.nn.Module):
    def __init__(self, kind, cfg, state_dict, lower_fn, out_channels=3):
        super().__init__()
        self.kind = kind
        self.cfg = cfg
        self._sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
        self._lower = lower_fn
        self.out_channels = out_channels
        self._engines = {}

    def engine_for(self, batch, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("diffpure_b200 runs on a B200 GPU only; got device %s (no CPU fallback)" % device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (int(batch), idx)
        eng = self._engines.get(key)
        if eng is None:
            prog = self._lower(self.cfg, self._sd, int(batch))
            eng = Engine(prog, device=idx)
            self._engines[key] = eng
        return eng

    def forward(self, x, t):
        """x: [B,3,H,W]; t: [B] conditioning exactly as the reference module receives it."""
        eng = self.engine_for(x.shape[0], x.device)
        return eng.unet_forward(x, t.float())

    def release(self):
        for e in self._engines.values():
            e.close()
        self._engines.clear()
