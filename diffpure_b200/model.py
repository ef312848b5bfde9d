"""Engine-backed score models: drop-in for the reference's UNet modules (`model(x, t) -> eps`).

The object is an nn.Module without parameters so that `.to()`, `.eval()` and nn.DataParallel wrapping of
the runners keep working (eval_sde_adv.py:227-229); the weights live in ONE packed device blob per GPU, shared by the
engines of every batch size (an engine = tensor maps + CUDA graphs + activation pool for one batch size). Engines are
kept in a small LRU (a ragged last batch or an attack with a varying batch size must not accumulate activation pools:
22 GB each for ADM at B=32); creation is serialised by a lock (nn.DataParallel calls replicas from threads).
"""
import threading
from collections import OrderedDict

import torch

from .engine import Engine, WeightBlob

MAX_ENGINES_PER_DEVICE = 3


class ScoreModel(torch.nn.Module):
    def __init__(self, kind, cfg, state_dict, lower_fn, out_channels=3, lower_vjp_fn=None):
        super().__init__()
        self.kind = kind
        self.cfg = cfg
        self._sd = {k: v.detach().float().cpu() for k, v in state_dict.items()}
        self._lower = lower_fn
        self._lower_vjp = lower_vjp_fn    # forward + data-gradient program (None: no backward for this network yet)
        self.out_channels = out_channels
        self._engines = OrderedDict()     # (batch, device index) -> Engine, most recently used last
        self._blobs = {}                  # device index -> WeightBlob
        self._lock = threading.RLock()

    def __deepcopy__(self, memo):         # nn.DataParallel.replicate / copy.deepcopy: replicas share the engines
        return self

    def blob_for(self, idx, program=None, tag="fwd"):
        with self._lock:
            blob = self._blobs.get((tag, idx))
            if blob is None:
                program = program or self._lower(self.cfg, self._sd, 1)
                blob = WeightBlob(program, idx)
                self._blobs[(tag, idx)] = blob
            return blob

    def engine_for(self, batch, device, vjp=False):
        """The engine of one batch size on one device; vjp=True: the forward + data-gradient program (`Engine.unet_vjp`)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("diffpure_b200 runs on a B200 GPU only; got device %s (no CPU fallback)" % device)
        if vjp and self._lower_vjp is None:
            raise NotImplementedError(f"diffpure_b200: the input-gradient program of the '{self.kind}' network is not "
                                      f"implemented (available: DDPM++ / 'ncsnpp')")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (("vjp", int(batch)) if vjp else int(batch), idx)
        with self._lock:
            eng = self._engines.get(key)
            if eng is not None:
                self._engines.move_to_end(key)
                return eng
            prog = (self._lower_vjp if vjp else self._lower)(self.cfg, self._sd, int(batch))
            same_dev = [k for k in self._engines if k[1] == idx]
            while len(same_dev) >= MAX_ENGINES_PER_DEVICE:      # evict the least recently used engine of this device
                self._engines.pop(same_dev.pop(0)).close()
            eng = Engine(prog, device=idx, blob=self.blob_for(idx, prog, "vjp" if vjp else "fwd"))
            self._engines[key] = eng
            return eng

    def adopt_engine(self, eng):
        """Register an engine built elsewhere (bench.py shares the one it already timed)."""
        with self._lock:
            self._engines[(eng.B, eng.device)] = eng
            self._blobs.setdefault(("fwd", eng.device), eng.blob)

    def forward(self, x, t):
        """x: [B,3,H,W]; t: [B] conditioning exactly as the reference module receives it."""
        eng = self.engine_for(x.shape[0], x.device)
        return eng.unet_forward(x, t.float())

    def release(self):
        with self._lock:
            for e in self._engines.values():
                e.close()
            self._engines.clear()
            self._blobs.clear()
