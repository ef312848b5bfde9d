"""ctypes binding of libdiffpure_b200.so (include/diffpure_b200.h). No fallback: a missing or broken
library raises -- the product path is the CUDA library."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiffpure_b200.so")

DP_UPDATE_LINEAR = 0
DP_UPDATE_LEARNED_RANGE = 1
DP_UPDATE_LINEAR_ANCHORED = 2


class EmbedDesc(C.Structure):
    _fields_ = [("out_bf16", C.c_void_p), ("B", C.c_int), ("dim", C.c_int), ("cos_first", C.c_int),
                ("half_minus_1", C.c_int)]


class GemmASeg(C.Structure):
    _fields_ = [("act_bf16", C.c_void_p), ("C", C.c_int), ("c_total", C.c_int), ("taps", C.c_int),
                ("stride", C.c_int), ("pad", C.c_int)]


class GemmDesc(C.Structure):
    _fields_ = [("a", GemmASeg * 2), ("nseg", C.c_int), ("w_bf16", C.c_void_p), ("w_rows", C.c_longlong),
                ("w_pitch", C.c_longlong), ("w_cols", C.c_longlong), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("N", C.c_int), ("batch", C.c_int), ("a_batch_rows", C.c_int), ("b_batch_rows", C.c_int),
                ("out_batch_stride", C.c_longlong), ("inner", C.c_int), ("a_inner_k", C.c_int), ("a_inner_rows", C.c_int),
                ("b_inner_k", C.c_int), ("b_inner_rows", C.c_int), ("out_inner_stride", C.c_longlong),
                ("bias", C.c_void_p), ("bias_along_m", C.c_int),
                ("rowvec", C.c_void_p), ("rowvec_ld", C.c_int), ("rowvec_rows_per_sample", C.c_int),
                ("rowscale", C.c_void_p), ("resid", C.c_void_p), ("alpha", C.c_float), ("silu", C.c_int),
                ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p), ("ldc", C.c_longlong), ("stats", C.c_void_p),
                ("softmax", C.c_int), ("softmax_scale", C.c_float), ("rowsum_out", C.c_void_p),
                ("gn_out_bf16", C.c_void_p), ("gn_gamma", C.c_void_p), ("gn_beta", C.c_void_p), ("gn_groups", C.c_int),
                ("gn_eps", C.c_float), ("gn_silu", C.c_int)]


class GnDesc(C.Structure):
    _fields_ = [("src0", C.c_void_p), ("stats0", C.c_void_p), ("C0", C.c_int), ("P0", C.c_int),
                ("src0_is_bf16", C.c_int), ("src1", C.c_void_p), ("stats1", C.c_void_p), ("C1", C.c_int), ("P1", C.c_int),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("film", C.c_void_p), ("film_ld", C.c_int),
                ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("groups", C.c_int), ("eps", C.c_float),
                ("silu", C.c_int), ("resample", C.c_int), ("out_bf16", C.c_void_p), ("raw_bf16", C.c_void_p),
                ("raw_f32", C.c_void_p)]


class StatsDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("B", C.c_int), ("HW", C.c_int), ("C", C.c_int), ("stats", C.c_void_p)]


class ConvInDesc(C.Structure):
    _fields_ = [("w", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p), ("stats", C.c_void_p),
                ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cout", C.c_int)]


class AttnSmallDesc(C.Structure):
    _fields_ = [("qkv_bf16", C.c_void_p), ("out_bf16", C.c_void_p), ("B", C.c_int), ("T", C.c_int),
                ("heads", C.c_int), ("d", C.c_int), ("scale", C.c_float)]


class SoftmaxDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("out_bf16", C.c_void_p), ("rows", C.c_longlong), ("T", C.c_int)]


class UpdateDesc(C.Structure):
    _fields_ = [("eps", C.c_void_p), ("ld", C.c_int), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cout", C.c_int)]


class GnBwdDesc(C.Structure):
    _fields_ = [("src0", C.c_void_p), ("stats0", C.c_void_p), ("C0", C.c_int), ("P0", C.c_int), ("src0_is_bf16", C.c_int),
                ("src1", C.c_void_p), ("stats1", C.c_void_p), ("C1", C.c_int), ("P1", C.c_int),
                ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("groups", C.c_int), ("eps", C.c_float),
                ("silu", C.c_int), ("resample", C.c_int), ("g", C.c_void_p), ("add0", C.c_void_p),
                ("add0_scale", C.c_float), ("add1", C.c_void_p), ("d0_f32", C.c_void_p), ("d0_bf16", C.c_void_p),
                ("d1_f32", C.c_void_p), ("film", C.c_void_p), ("film_ld", C.c_int)]


class SoftmaxBwdDesc(C.Structure):
    _fields_ = [("pnum_bf16", C.c_void_p), ("rowsum", C.c_void_p), ("dp", C.c_void_p), ("ds_bf16", C.c_void_p),
                ("pn_bf16", C.c_void_p), ("rows", C.c_longlong), ("T", C.c_int)]


class TransposeDesc(C.Structure):
    _fields_ = [("in_bf16", C.c_void_p), ("out_bf16", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int),
                ("ld_in", C.c_int), ("ld_out", C.c_int), ("batch", C.c_int), ("in_batch_stride", C.c_longlong),
                ("out_batch_stride", C.c_longlong)]


class AttnSmallBwdDesc(C.Structure):
    _fields_ = [("qkv_bf16", C.c_void_p), ("go_bf16", C.c_void_p), ("out_bf16", C.c_void_p), ("B", C.c_int),
                ("T", C.c_int), ("heads", C.c_int), ("d", C.c_int), ("scale", C.c_float)]


class GradInDesc(C.Structure):
    _fields_ = [("out_bf16", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("C", C.c_int),
                ("Cpad", C.c_int)]


class PadInDesc(C.Structure):
    _fields_ = [("out_bf16", C.c_void_p), ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Cpad", C.c_int)]


class AttnBlockDesc(C.Structure):
    _fields_ = [("hn_bf16", C.c_void_p), ("w_bf16", C.c_void_p), ("bias", C.c_void_p), ("resid", C.c_void_p),
                ("out_f32", C.c_void_p), ("stats", C.c_void_p), ("B", C.c_int), ("T", C.c_int), ("C", C.c_int),
                ("scale", C.c_float), ("alpha", C.c_float)]


class PurifyParams(C.Structure):
    _fields_ = [("steps", C.c_int), ("update_kind", C.c_int), ("ncoef", C.c_int), ("cond", C.c_void_p),
                ("coef", C.c_void_p), ("init_scale_x", C.c_float), ("init_scale_e", C.c_float),
                ("init_noise", C.c_void_p), ("step_noise", C.c_void_p), ("seed", C.c_uint64),
                ("sample_offset", C.c_uint64), ("anchor", C.c_void_p), ("states", C.c_void_p),
                ("in_h", C.c_int), ("in_w", C.c_int), ("in_unit_range", C.c_int), ("out_h", C.c_int), ("out_w", C.c_int),
                ("out_unit_range", C.c_int), ("out_mean", C.c_float * 3), ("out_std", C.c_float * 3)]


# every symbol include/diffpure_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "dp_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "dp_destroy": (None, [C.c_void_p]),
    "dp_last_error": (C.c_char_p, [C.c_void_p]),
    "dp_version": (C.c_int, []),
    "dp_device_sm_count": (C.c_int, [C.c_void_p]),
    "dp_buffer_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
    "dp_buffer_adopt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
    "dp_buffer_ptr": (C.c_void_p, [C.c_void_p, C.c_int]),
    "dp_buffer_write": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]),
    "dp_buffer_read": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]),
    "dp_bytes_allocated": (C.c_size_t, [C.c_void_p]),
    "dp_op_embed": (C.c_int, [C.c_void_p, C.POINTER(EmbedDesc)]),
    "dp_op_gemm": (C.c_int, [C.c_void_p, C.POINTER(GemmDesc)]),
    "dp_op_gn_apply": (C.c_int, [C.c_void_p, C.POINTER(GnDesc)]),
    "dp_op_stats": (C.c_int, [C.c_void_p, C.POINTER(StatsDesc)]),
    "dp_op_conv_in": (C.c_int, [C.c_void_p, C.POINTER(ConvInDesc)]),
    "dp_op_attn_small": (C.c_int, [C.c_void_p, C.POINTER(AttnSmallDesc)]),
    "dp_op_attn_block": (C.c_int, [C.c_void_p, C.POINTER(AttnBlockDesc)]),
    "dp_op_softmax_rows": (C.c_int, [C.c_void_p, C.POINTER(SoftmaxDesc)]),
    "dp_op_update": (C.c_int, [C.c_void_p, C.POINTER(UpdateDesc)]),
    "dp_op_gn_bwd": (C.c_int, [C.c_void_p, C.POINTER(GnBwdDesc)]),
    "dp_op_softmax_bwd": (C.c_int, [C.c_void_p, C.POINTER(SoftmaxBwdDesc)]),
    "dp_op_transpose": (C.c_int, [C.c_void_p, C.POINTER(TransposeDesc)]),
    "dp_op_attn_small_bwd": (C.c_int, [C.c_void_p, C.POINTER(AttnSmallBwdDesc)]),
    "dp_op_grad_in": (C.c_int, [C.c_void_p, C.POINTER(GradInDesc)]),
    "dp_op_pad_in": (C.c_int, [C.c_void_p, C.POINTER(PadInDesc)]),
    "dp_unet_vjp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dp_program_size": (C.c_int, [C.c_void_p]),
    "dp_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "dp_unet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dp_purify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PurifyParams), C.c_void_p]),
    "dp_launches_per_eval": (C.c_int, [C.c_void_p]),
    "dp_launches_per_step": (C.c_int, [C.c_void_p]),
    "dp_gemm_pair_count": (C.c_int, [C.c_void_p]),
    "dp_gemm_fused_gn_count": (C.c_int, [C.c_void_p]),
    "dp_profile_ops": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int),
                                 C.POINTER(C.c_double), C.c_int]),
    "dp_normal_host": (C.c_float, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every prototype. Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C diffpure_b200/csrc`). diffpure_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DPError(RuntimeError):
    pass


def check(lib, handle, rc, what):
    if rc != 0:
        msg = lib.dp_last_error(handle)
        raise DPError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
