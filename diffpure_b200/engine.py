"""Engine backend: materialises a `Program` through the C ABI and runs it.

PyTorch is used only for device memory of the caller-visible tensors and for streams; every kernel on the
path lives in libdiffpure_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as _lib
from .program import Program, View


def _align(n, a=256):
    return (n + a - 1) // a * a


def _to_bytes(t, dtype):
    t = t.detach().cpu().contiguous()
    if dtype == "bf16":
        return t.to(torch.bfloat16).view(torch.int16).numpy().tobytes()
    return t.to(torch.float32).numpy().tobytes()


class WeightBlob:
    """All constants of a lowered model (bf16 weights, fp32 biases / norm parameters / tables) packed into ONE device
    allocation: one upload, one NCCL broadcast, shared by the engines of every batch size on that device.

    The memory is a torch uint8 CUDA tensor (PyTorch = device memory + collectives only); engines adopt its pointer
    through dp_buffer_adopt. `layout` maps constant name -> (byte offset, byte size)."""

    def __init__(self, program: Program, device, upload=True):
        consts = [t for t in program.tensors if t.init is not None]
        self.layout, total = {}, 0
        for t in consts:
            self.layout[t.name] = (total, t.nbytes)
            total += _align(t.nbytes)
        self.nbytes = total
        # ("cpu" is accepted for the host-logic tests of packing / broadcast; engines only adopt CUDA blobs)
        self.device = torch.device(device) if isinstance(device, (str, torch.device)) else torch.device("cuda", int(device))
        self.data = torch.empty(max(total, 256), dtype=torch.uint8, device=self.device)
        assert self.data.data_ptr() % 256 == 0 or self.device.type == "cpu"
        if upload:
            host = torch.zeros(max(total, 256), dtype=torch.uint8)
            if self.device.type == "cuda":
                host = host.pin_memory()
            for t in consts:
                off, n = self.layout[t.name]
                src = t.init.detach().cpu().contiguous()
                src = src.to(torch.bfloat16) if t.dtype == "bf16" else src.to(torch.float32)
                host[off:off + n] = src.view(torch.uint8).reshape(-1)
            self.data.copy_(host, non_blocking=False)

    def matches(self, program: Program):
        consts = [t for t in program.tensors if t.init is not None]
        return len(consts) == len(self.layout) and all(
            t.name in self.layout and self.layout[t.name][1] == t.nbytes for t in consts)

    def broadcast(self, src=0):
        """One collective for all weights: the packed blob itself (NCCL; rank `src` holds the real values)."""
        import torch.distributed as dist
        dist.broadcast(self.data, src=src)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()

    def tensor(self, name, dtype):
        """Host copy of one constant (tests)."""
        off, n = self.layout[name]
        raw = self.data[off:off + n].cpu()
        return raw.view(torch.bfloat16).float() if dtype == "bf16" else raw.view(torch.float32)


class Engine:
    def __init__(self, program: Program, device=0, pool=True, blob: WeightBlob = None):
        self.lib = _lib.load()
        self.program = program
        self.device = int(device)
        self.B, self.H, self.W = program.B, program.H, program.W
        self.out_channels = program.meta.get("out_channels", 3)
        h = C.c_void_p()
        rc = self.lib.dp_create(C.byref(h), self.device)
        if rc != 0:
            raise _lib.DPError(f"dp_create failed (code {rc}): {self.lib.dp_last_error(None).decode()}")
        self.h = h
        self._ptr = {}       # tensor index -> device address
        self._loc = {}       # tensor index -> (buffer id, byte offset)
        self.const_bytes = 0
        self.act_bytes = 0
        try:
            with torch.cuda.device(self.device):
                self.blob = blob if blob is not None else WeightBlob(program, self.device)
                self._adopt_constants()
                self._place_activations(pool)
                self._emit_ops()
                self._check(self.lib.dp_finalize(self.h, self.B, self.H, self.W), "dp_finalize")
        except Exception:
            self.close()
            raise

    # ------------------------------------------------------------------------------------------
    def _check(self, rc, what):
        _lib.check(self.lib, self.h, rc, what)

    def _alloc(self, nbytes):
        bid = C.c_int()
        self._check(self.lib.dp_buffer_alloc(self.h, nbytes, C.byref(bid)), "dp_buffer_alloc")
        return bid.value, self.lib.dp_buffer_ptr(self.h, bid.value)

    def _adopt_constants(self):
        """The program's constants live in the (possibly shared) weight blob; the engine only records their addresses."""
        blob = self.blob
        if blob.device.type != "cuda" or blob.device.index != self.device:
            raise ValueError("weight blob lives on another device")
        if not blob.matches(self.program):
            raise ValueError("weight blob layout does not match this program's constants")
        self.const_bytes = blob.nbytes
        if blob.nbytes == 0:
            return
        bid = C.c_int()
        self._check(self.lib.dp_buffer_adopt(self.h, blob.data.data_ptr(), blob.nbytes, C.byref(bid)), "dp_buffer_adopt")
        base = blob.data.data_ptr()
        self.weights_buffer = (bid.value, base, blob.nbytes)
        for t in self.program.tensors:
            if t.init is not None:
                off = blob.layout[t.name][0]
                self._ptr[t.index] = base + off
                self._loc[t.index] = (bid.value, off)

    def _place_activations(self, pool):
        prog = self.program
        first, last = prog.first_use(), prog.last_use()
        acts = [t for t in prog.tensors if t.init is None and t.index in first]
        by_first = {}
        for t in acts:
            by_first.setdefault(first[t.index], []).append(t)
        by_last = {}
        for t in acts:
            by_last.setdefault(last[t.index], []).append(t)
        free = []      # (nbytes, ptr)
        owner = {}     # tensor index -> (nbytes, ptr)
        for i in range(len(prog.ops)):
            for t in by_first.get(i, []):
                need = _align(t.nbytes)
                pick = None
                if pool:
                    cands = [f for f in free if need <= f[0] <= 2 * need]
                    if cands:
                        pick = min(cands)
                        free.remove(pick)
                if pick is None:
                    bid, ptr = self._alloc(need)
                    pick = (need, ptr, bid)
                    self.act_bytes += need
                owner[t.index] = pick
                self._ptr[t.index] = pick[1]
                self._loc[t.index] = (pick[2], 0)
            for t in by_last.get(i, []):
                free.append(owner[t.index])

    def _p(self, v, elem_bytes=None):
        if v is None:
            return None
        assert isinstance(v, View)
        eb = 4 if v.tensor.dtype == "f32" else 2
        return self._ptr[v.tensor.index] + v.offset * eb

    def _emit_ops(self):
        L = self.lib
        for op in self.program.ops:
            a = op.args
            if op.kind == "embed":
                d = _lib.EmbedDesc(self._p(a["out"]), a["B"], a["dim"], a["cos_first"], a["half_minus_1"])
                self._check(L.dp_op_embed(self.h, C.byref(d)), "dp_op_embed")
            elif op.kind == "gemm":
                d = _lib.GemmDesc()
                for i, seg in enumerate(a["a"]):
                    d.a[i] = _lib.GemmASeg(self._p(seg.act), seg.C, seg.c_total, seg.taps, seg.stride, seg.pad)
                d.nseg = len(a["a"])
                d.w_bf16 = self._p(a["w"]); d.w_rows = a["w_rows"]; d.w_pitch = a["w_pitch"]
                d.w_cols = a["w_cols"]
                d.inner, d.a_inner_k, d.b_inner_k = a["inner"], a["a_inner_k"], a["b_inner_k"]
                d.a_inner_rows = a["a_inner_rows"]
                d.b_inner_rows, d.out_inner_stride = a["b_inner_rows"], a["out_inner_stride"]
                d.B, d.H, d.W, d.N = a["B"], a["H"], a["W"], a["N"]
                d.batch, d.a_batch_rows, d.b_batch_rows = a["batch"], a["a_batch_rows"], a["b_batch_rows"]
                d.out_batch_stride = a["out_batch_stride"]
                d.bias = self._p(a["bias"]); d.bias_along_m = a["bias_along_m"]
                d.rowvec = self._p(a["rowvec"]); d.rowvec_ld = a["rowvec_ld"]
                d.rowvec_rows_per_sample = a["rowvec_rows_per_sample"]
                d.rowscale = self._p(a["rowscale"]); d.resid = self._p(a["resid"])
                d.alpha = a["alpha"]; d.silu = a["silu"]
                d.out_f32 = self._p(a["out_f32"]); d.out_bf16 = self._p(a["out_bf16"]); d.ldc = a["ldc"]
                d.stats = self._p(a["stats"])
                d.softmax = a["softmax"]; d.softmax_scale = a["softmax_scale"]; d.rowsum_out = self._p(a["rowsum_out"])
                d.gn_out_bf16 = self._p(a["gn_out"]); d.gn_gamma = self._p(a["gn_gamma"])
                d.gn_beta = self._p(a["gn_beta"]); d.gn_groups = a["gn_groups"]; d.gn_eps = a["gn_eps"]
                d.gn_silu = a["gn_silu"]
                self._check(L.dp_op_gemm(self.h, C.byref(d)), "dp_op_gemm")
            elif op.kind == "gn_apply":
                d = _lib.GnDesc(self._p(a["src0"]), self._p(a["stats0"]), a["C0"], a["P0"],
                                1 if a["src0"].tensor.dtype == "bf16" else 0, self._p(a["src1"]),
                                self._p(a["stats1"]), a["C1"], a["P1"], self._p(a["gamma"]), self._p(a["beta"]),
                                self._p(a["film"]), a["film_ld"], a["B"], a["H"], a["W"], a["groups"], a["eps"],
                                a["silu"], a["resample"], self._p(a["out_bf16"]), self._p(a["raw_bf16"]),
                                self._p(a["raw_f32"]))
                self._check(L.dp_op_gn_apply(self.h, C.byref(d)), "dp_op_gn_apply")
            elif op.kind == "conv_in":
                d = _lib.ConvInDesc(self._p(a["w"]), self._p(a["bias"]), self._p(a["out"]), self._p(a["stats"]),
                                    a["B"], a["H"], a["W"], a["Cout"])
                self._check(L.dp_op_conv_in(self.h, C.byref(d)), "dp_op_conv_in")
            elif op.kind == "attn_small":
                d = _lib.AttnSmallDesc(self._p(a["qkv"]), self._p(a["out"]), a["B"], a["T"], a["heads"], a["d"],
                                       a["scale"])
                self._check(L.dp_op_attn_small(self.h, C.byref(d)), "dp_op_attn_small")
            elif op.kind == "attn_block":
                d = _lib.AttnBlockDesc(self._p(a["hn"]), self._p(a["w"]), self._p(a["bias"]), self._p(a["resid"]),
                                       self._p(a["out_f32"]), self._p(a["stats"]), a["B"], a["T"], a["C"], a["scale"],
                                       a["alpha"])
                self._check(L.dp_op_attn_block(self.h, C.byref(d)), "dp_op_attn_block")
            elif op.kind == "update":
                d = _lib.UpdateDesc(self._p(a["eps"]), a["ld"], a["B"], a["H"], a["W"], a["Cout"])
                self._check(L.dp_op_update(self.h, C.byref(d)), "dp_op_update")
            elif op.kind == "softmax_rows":
                d = _lib.SoftmaxDesc(self._p(a["src"]), self._p(a["out"]), a["rows"], a["T"])
                self._check(L.dp_op_softmax_rows(self.h, C.byref(d)), "dp_op_softmax_rows")
            elif op.kind == "gn_bwd":
                d = _lib.GnBwdDesc(self._p(a["src0"]), self._p(a["stats0"]), a["C0"], a["P0"],
                                   1 if a["src0"].tensor.dtype == "bf16" else 0, self._p(a["src1"]),
                                   self._p(a["stats1"]), a["C1"], a["P1"], self._p(a["gamma"]), self._p(a["beta"]),
                                   a["B"], a["H"], a["W"], a["groups"], a["eps"], a["silu"], a["resample"],
                                   self._p(a["g"]), self._p(a["add0"]), a["add0_scale"], self._p(a["add1"]),
                                   self._p(a["d0_f32"]), self._p(a["d0_bf16"]), self._p(a["d1_f32"]),
                                   self._p(a.get("film")), a.get("film_ld", 0))
                self._check(L.dp_op_gn_bwd(self.h, C.byref(d)), "dp_op_gn_bwd")
            elif op.kind == "softmax_bwd":
                d = _lib.SoftmaxBwdDesc(self._p(a["pnum"]), self._p(a["rowsum"]), self._p(a["dp"]), self._p(a["ds"]),
                                        self._p(a["pn"]), a["rows"], a["T"])
                self._check(L.dp_op_softmax_bwd(self.h, C.byref(d)), "dp_op_softmax_bwd")
            elif op.kind == "transpose":
                d = _lib.TransposeDesc(self._p(a["src"]), self._p(a["out"]), a["rows"], a["cols"], a["ld_in"],
                                       a["ld_out"], a["batch"], a["in_batch_stride"], a["out_batch_stride"])
                self._check(L.dp_op_transpose(self.h, C.byref(d)), "dp_op_transpose")
            elif op.kind == "attn_small_bwd":
                d = _lib.AttnSmallBwdDesc(self._p(a["qkv"]), self._p(a["go"]), self._p(a["out"]), a["B"], a["T"],
                                          a["heads"], a["d"], a["scale"])
                self._check(L.dp_op_attn_small_bwd(self.h, C.byref(d)), "dp_op_attn_small_bwd")
            elif op.kind == "pad_in":
                d = _lib.PadInDesc(self._p(a["out"]), a["B"], a["H"], a["W"], a["Cpad"])
                self._check(L.dp_op_pad_in(self.h, C.byref(d)), "dp_op_pad_in")
            elif op.kind == "grad_in":
                d = _lib.GradInDesc(self._p(a["out"]), a["B"], a["H"], a["W"], a["C"], a["Cpad"])
                self._check(L.dp_op_grad_in(self.h, C.byref(d)), "dp_op_grad_in")
            else:
                raise ValueError(f"unknown op kind {op.kind}")

    # ------------------------------------------------------------------------------------------
    @property
    def launches_per_eval(self):
        return self.lib.dp_launches_per_eval(self.h)

    @property
    def launches_per_step(self):
        """Kernels per loop step (the update is the output conv's epilogue: no extra launch)."""
        return self.lib.dp_launches_per_step(self.h)

    @property
    def pair_gemms(self):
        """GEMM ops that run on CTA-pair (cta_group::2) tiles."""
        return self.lib.dp_gemm_pair_count(self.h)

    @property
    def fused_gn_gemms(self):
        """GEMM ops whose epilogue applies GroupNorm(+SiLU) with the sample's accumulators resident in TMEM."""
        return self.lib.dp_gemm_fused_gn_count(self.h)

    def _dev(self):
        return torch.device("cuda", self.device)

    def _stream(self):
        st = torch.cuda.current_stream(self._dev()).cuda_stream
        return C.c_void_p(st) if st else None

    def _prep(self, x, hw=None):
        want = (self.B, 3) + (tuple(hw) if hw else (self.H, self.W))
        assert tuple(x.shape) == want, (tuple(x.shape), want)
        return x.to(device=self._dev(), dtype=torch.float32).contiguous()

    def unet_forward(self, x, cond):
        """x: [B,3,H,W]; cond: [B] float. Returns the UNet output [B,Cout,H,W] (fp32, on the engine's device)."""
        x = self._prep(x)
        cond = cond.to(device=self._dev(), dtype=torch.float32).contiguous()
        assert cond.shape == (self.B,)
        out = torch.empty((self.B, self.out_channels, self.H, self.W), device=self._dev(), dtype=torch.float32)
        # enqueued on torch's current stream (stream-ordered with the caller's tensors); the default stream maps to
        # the blocking NULL-stream contract of the C ABI
        self._check(self.lib.dp_unet_forward(self.h, x.data_ptr(), cond.data_ptr(), out.data_ptr(), self._stream()),
                    "dp_unet_forward")
        return out

    def unet_vjp(self, x, cond, g_out):
        """J(x, cond)^T g_out of one UNet evaluation (program built by a `lower_vjp`): [B,3,H,W] fp32."""
        x = self._prep(x)
        cond = cond.to(device=self._dev(), dtype=torch.float32).contiguous()
        g_out = g_out.to(device=self._dev(), dtype=torch.float32).contiguous()
        assert cond.shape == (self.B,) and g_out.shape[0] == self.B and g_out.shape[2:] == x.shape[2:]
        out = torch.empty((self.B, 3, self.H, self.W), device=self._dev(), dtype=torch.float32)
        self._keep = (x, cond, g_out)
        self._check(self.lib.dp_unet_vjp(self.h, x.data_ptr(), cond.data_ptr(), g_out.data_ptr(), out.data_ptr(),
                                         self._stream()), "dp_unet_vjp")
        return out

    def purify(self, x0, cond, coef, init_scale_x, init_scale_e, *, update_kind=_lib.DP_UPDATE_LINEAR,
               init_noise=None, step_noise=None, seed=0, sample_offset=0, anchor=None, states=None,
               in_unit_range=False, out_hw=None, out_unit_range=False, out_norm=None):
        """Runs the whole loop on the device. cond: [steps] host floats; coef: [steps, ncoef] host floats.
        states: optional [steps+1,B,3,H,W] fp32 device tensor receiving the state before every step and the final one.
        Fused pre / post steps (eval_sde_adv.py:73-89): x0 may have another spatial size (bilinear resize to the model
        grid) and be in [0,1] (in_unit_range); out_hw resizes the result, out_unit_range maps it to [0,1], out_norm =
        (mean[3], std[3]) applies the classifier's normalisation."""
        in_hw = tuple(x0.shape[2:]) if tuple(x0.shape[2:]) != (self.H, self.W) else None
        x0 = self._prep(x0, in_hw)
        cond = np.ascontiguousarray(np.asarray(cond, dtype=np.float32))
        coef = np.ascontiguousarray(np.asarray(coef, dtype=np.float32))
        steps = cond.shape[0]
        assert coef.ndim == 2 and coef.shape[0] == steps
        model_shape = (self.B, 3, self.H, self.W)
        if init_noise is not None:
            init_noise = init_noise.to(device=self._dev(), dtype=torch.float32).contiguous()
            assert tuple(init_noise.shape) == model_shape
        if step_noise is not None:
            step_noise = step_noise.to(device=self._dev(), dtype=torch.float32).contiguous()
            assert tuple(step_noise.shape) == (steps,) + model_shape
        if anchor is not None:
            anchor = self._prep(anchor)
        if states is not None:
            assert states.is_cuda and states.dtype == torch.float32 and states.is_contiguous() and \
                tuple(states.shape) == (steps + 1,) + model_shape
        out = torch.empty((self.B, 3) + (tuple(out_hw) if out_hw else (self.H, self.W)), device=self._dev(),
                          dtype=torch.float32)
        mean, std = out_norm if out_norm is not None else ((0.0,) * 3, (0.0,) * 3)
        p = _lib.PurifyParams(steps, update_kind, coef.shape[1], cond.ctypes.data, coef.ctypes.data,
                              float(init_scale_x), float(init_scale_e),
                              init_noise.data_ptr() if init_noise is not None else None,
                              step_noise.data_ptr() if step_noise is not None else None, int(seed),
                              int(sample_offset), anchor.data_ptr() if anchor is not None else None,
                              states.data_ptr() if states is not None else None,
                              in_hw[0] if in_hw else 0, in_hw[1] if in_hw else 0, 1 if in_unit_range else 0,
                              out_hw[0] if out_hw else 0, out_hw[1] if out_hw else 0, 1 if out_unit_range else 0,
                              (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std]))
        self._keep = (x0, init_noise, step_noise, anchor)    # inputs stay alive until the enqueued work has run
        self._check(self.lib.dp_purify(self.h, x0.data_ptr(), out.data_ptr(), C.byref(p), self._stream()), "dp_purify")
        return out

    OP_KINDS = ("embed", "gemm", "gn_apply", "stats", "stats_reduce", "conv_in", "attn_small",
                "softmax_rows", "update", "gn_bwd", "softmax_bwd", "transpose", "attn_small_bwd", "grad_in", "gn_finalize", "pad_in", "attn_block")

    def profile_ops(self, mode=0):
        """Per-op device time (ms), kind and executed GEMM flops of one eagerly-run UNet evaluation."""
        n = self.launches_per_eval
        ms = (C.c_float * n)()
        kinds = (C.c_int * n)()
        flops = (C.c_double * n)()
        self._check(self.lib.dp_profile_ops(self.h, mode, ms, kinds, flops, n), "dp_profile_ops")
        return [(self.OP_KINDS[kinds[i]], float(ms[i]), float(flops[i])) for i in range(n)]

    def read_tensor(self, name):
        """Debug: copy an engine tensor back to the host (intermediate values need pool=False)."""
        t = next(t for t in self.program.tensors if t.name == name)
        bid, off = self._loc[t.index]
        host = np.empty(t.numel, dtype=np.int16 if t.dtype == "bf16" else np.float32)
        self._check(self.lib.dp_buffer_read(self.h, bid, off, host.ctypes.data, host.nbytes), "dp_buffer_read")
        out = torch.from_numpy(host)
        return out.view(torch.bfloat16).float() if t.dtype == "bf16" else out

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.dp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
