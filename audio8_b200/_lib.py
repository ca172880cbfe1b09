"""ctypes binding of libaudio8_b200.so (the C ABI declared in include/audio8_b200.h).

There is no fallback: if the shared object is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_TAG = os.environ.get("A8_LIB_TAG", "")  # kernel experiments: load libaudio8_b200_<tag>.so (see build.py)
LIB_PATH = os.path.join(HERE, "libaudio8_b200" + ("_" + _TAG if _TAG else "") + ".so")

MAJOR_K, MAJOR_MN = 0, 1
OUT_BF16, OUT_F32, OUT_F32_ATOMIC = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_GELU_DZ = 0, 1, 2
AUX_NONE, AUX_ADD, AUX_MUL_GELU_GRAD, AUX_MUL = 0, 1, 2, 3

_i32x4 = C.c_int32 * 4


class Operand(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("dims", C.c_int64 * 4), ("strides", C.c_int64 * 3), ("major", C.c_int32),
                ("base", _i32x4), ("ck", _i32x4), ("cb", _i32x4), ("cr", _i32x4), ("cl", _i32x4), ("ch", _i32x4)]


class Gemm(C.Structure):
    _fields_ = [("a", Operand), ("b", Operand), ("M", C.c_int32), ("N", C.c_int32), ("lo_count", C.c_int32),
                ("hi_count", C.c_int32), ("k_blocks", C.c_int32), ("k_inner", C.c_int32), ("split_k", C.c_int32),
                ("block_n", C.c_int32), ("c", C.c_void_p), ("c_dtype", C.c_int32), ("act", C.c_int32),
                ("ldc", C.c_int64), ("c_stride_lo", C.c_int64), ("c_stride_hi", C.c_int64), ("z_out", C.c_void_p),
                ("aux", C.c_void_p), ("aux_mode", C.c_int32), ("bias_stride_lo", C.c_int32), ("bias", C.c_void_p),
                ("alpha", C.c_float), ("reserved", C.c_int32), ("colsum", C.c_void_p)]


class A8Error(RuntimeError):
    pass


_P, _I, _L, _F, _Z, _D = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t, C.c_double

# name -> (restype, argtypes); every symbol declared in include/audio8_b200.h must be listed here
SIGNATURES = {
    "a8_version": (_I, []),
    "a8_last_error": (C.c_char_p, []),
    "a8_launch_count": (_L, []),
    "a8_launch_count_add": (_I, [_L]),
    "a8_set_seed_source": (_I, [_P]),
    "a8_gemm": (_I, [C.POINTER(Gemm), _P]),
    "a8_gemm_group": (_I, [C.POINTER(Gemm), _I, _P]),
    "a8_gemm_group_blob_bytes": (_Z, []),
    "a8_gemm_group_prepare": (_I, [C.POINTER(Gemm), _I, _P]),
    "a8_gemm_group_launch": (_I, [_P, _P]),
    "a8_ctc_scratch_floats": (_Z, [_I, _I, _I]),
    "a8_ctc_greedy": (_I, [_P, _L, _L, _L, _I, _I, _I, _P, _I, _P, _P, _P]),
    "a8_ctc_prep": (_I, [_P, _L, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "a8_ctc_forward": (_I, [_P, _L, _L, _L, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "a8_ctc_backward": (_I, [_P, _L, _L, _L, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _L, _I, _I, _P, _L, _L, _P]),
}

_U = C.c_uint64
SIGNATURES.update({
    "a8_layernorm_fwd": (_I, [_P, _P, _F, _U, _P, _P, _P, _F, _P, _P, _F, _U, _P, _P, _I, _I, _P]),
    "a8_layernorm_bwd": (_I, [_P, _P, _F, _U, _P, _P, _P, _P, _P, _P, _F, _U, _P, _P, _P, _I, _I, _P]),
    "a8_attn_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _F, _U, _P]),
    "a8_attn_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _U, _P]),
    "a8_attn_dropmask": (_I, [_P, _I, _I, _I, _F, _U, _P]),
    "a8_colsum": (_I, [_P, _L, _I, _I, _P, _P]),
    "a8_dropout": (_I, [_P, _P, _I, _L, _F, _U, _P]),
    "a8_gelu_bwd": (_I, [_P, _P, _P, _L, _P]),
    "a8_mul_dgelu": (_I, [_P, _P, _P, _L, _P]),
    "a8_log_softmax_fwd": (_I, [_P, _P, _I, _I, _P]),
    "a8_log_softmax_bwd": (_I, [_P, _L, _L, _L, _I, _P, _P, _I, _I, _P]),
    "a8_conv0_stats": (_I, [_P, _I, _L, _P, _I, _I, _I, _F, _P, _P, _P, _P]),
    "a8_conv0_fwd": (_I, [_P, _I, _L, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "a8_conv0_bwd": (_I, [_P, _I, _L, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "a8_rows_copy": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _P]),
    "a8_rows_set": (_I, [_P, _P, _I, _I, _P, _P]),
    "a8_rows_set_bwd": (_I, [_P, _P, _I, _I, _P, _P]),
    "a8_mask_apply": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "a8_cast": (_I, [_P, _I, _P, _I, _L, _P]),
    "a8_split3": (_I, [_P, _P, _I, _I, _I, _P]),
    "a8_vq_fwd": (_I, [_P, _P, _F, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "a8_vq_bwd": (_I, [_P, _P, _F, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "a8_contrastive_fwd": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "a8_cast_multi": (_I, [_P, _I, _P]),
    "a8_conv_pack": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "a8_conv_unpack": (_I, [_P, _I, _I, _I, _P, _P]),
    "a8_gemm_set_trace": (None, [_P]),
    "a8_posconv_norm_scratch_floats": (_L, [_I, _I, _I]),
    "a8_posconv_pack": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "a8_posconv_wn_bwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "a8_contrastive_bwd": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "a8_optim_grad_sqnorm": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "a8_optim_adamw": (_I, [_P, _P, _P, _I, _I, _P, _F, _F, _D, _D, _D, _D, _D, _D, _D, _I, _P, _P]),
    "a8_allreduce_mc": (_I, [_P, _L, _L, _I, _I, _F, _I, _P]),
    "a8_span_mask_draw": (_I, [_U, _P, _I, _I, _D, _I, _I, _P, _P, _P]),
    "a8_negatives_draw": (_I, [_U, _P, _P, _I, _I, _I, _P, _P]),
})

_lib = None


def load():
    """Load the shared object (once).  Raises if it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise A8Error(f"{LIB_PATH} not found — run `python -m audio8_b200.build` (no fallback path exists)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        raise A8Error(f"{what} failed ({rc}): {load().a8_last_error().decode()}")
