"""Python face of the C ABI: tensor-level wrappers around libaudio8_b200.so.

Every function here takes torch CUDA tensors, checks them, and enqueues hand-written sm_100a kernels on the
current CUDA stream through `ctypes`.  There is no CPU or PyTorch fallback: a missing library or a failed
launch raises.  (`tests/emu.py` contains a pure-PyTorch emulation of the same ABI that CPU tests inject to
check the host-side orchestration without a GPU; the product never imports it.)
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_NONE, AUX_ADD, AUX_MUL_GELU_GRAD, AUX_NONE, MAJOR_K, MAJOR_MN, OUT_BF16, OUT_F32,
                   OUT_F32_ATOMIC)

__all__ = ["Op", "GemmSpec", "gemm", "backend", "set_backend"]


class Op:
    """One GEMM operand: a bf16 buffer seen as a 4-D strided view plus the affine tile-coordinate map
    (include/audio8_b200.h: a8_operand_t).  `t` supplies the base address; `offset` is in elements."""

    __slots__ = ("t", "offset", "dims", "strides", "major", "base", "ck", "cb", "cr", "cl", "ch")

    def __init__(self, t, dims, strides, major, offset=0, base=(0, 0, 0, 0), ck=(0, 0, 0, 0), cb=(0, 0, 0, 0),
                 cr=(0, 0, 0, 0), cl=(0, 0, 0, 0), ch=(0, 0, 0, 0)):
        dims = list(dims) + [1] * (4 - len(dims))
        strides = list(strides)
        while len(strides) < 3:  # strides of size-1 dims are never used; keep them legal for the tensor map
            strides.append(max(strides[-1] * max(dims[len(strides)], 1), 8) if strides else 8)
        self.t, self.offset, self.dims, self.strides, self.major = t, offset, dims, strides, major
        self.base, self.ck, self.cb, self.cr, self.cl, self.ch = (tuple(v) for v in (base, ck, cb, cr, cl, ch))


class GemmSpec:
    """C[hi][lo][m][n] = epilogue(alpha * sum_k A[m][k] B[n][k])  (include/audio8_b200.h: a8_gemm_t)."""

    def __init__(self, a, b, M, N, k_blocks, c, ldc, c_dtype=OUT_BF16, lo_count=1, hi_count=1, k_inner=None,
                 split_k=1, block_n=0, c_offset=0, c_stride_lo=0, c_stride_hi=0, act=ACT_NONE, z_out=None,
                 aux=None, aux_mode=AUX_NONE, bias=None, bias_stride_lo=0, alpha=1.0):
        self.a, self.b, self.M, self.N, self.k_blocks = a, b, M, N, k_blocks
        self.k_inner = k_inner if k_inner is not None else k_blocks
        self.c, self.ldc, self.c_dtype, self.c_offset = c, ldc, c_dtype, c_offset
        self.lo_count, self.hi_count, self.split_k, self.block_n = lo_count, hi_count, split_k, block_n
        self.c_stride_lo, self.c_stride_hi = c_stride_lo, c_stride_hi
        self.act, self.z_out, self.aux, self.aux_mode = act, z_out, aux, aux_mode
        self.bias, self.bias_stride_lo, self.alpha = bias, bias_stride_lo, alpha


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, offset_elems=0):
    if t is None:
        return None
    return C.c_void_p(t.data_ptr() + offset_elems * t.element_size())


class CudaBackend:
    """Calls into libaudio8_b200.so.  Constructed lazily; raises if the library is absent."""

    name = "cuda"

    def __init__(self):
        self.lib = _lib.load()

    # ------------------------------------------------------------------ gemm
    @staticmethod
    def _operand(o):
        assert o.t.is_cuda and o.t.dtype == torch.bfloat16, "GEMM operands must be CUDA bf16 tensors"
        s = _lib.Operand()
        s.ptr = o.t.data_ptr() + 2 * o.offset
        for i in range(4):
            s.dims[i] = o.dims[i]
            s.base[i], s.ck[i], s.cb[i], s.cr[i], s.cl[i], s.ch[i] = o.base[i], o.ck[i], o.cb[i], o.cr[i], o.cl[i], o.ch[i]
        for i in range(3):
            s.strides[i] = o.strides[i]
        s.major = o.major
        return s

    def gemm(self, g):
        s = _lib.Gemm()
        s.a, s.b = self._operand(g.a), self._operand(g.b)
        s.M, s.N, s.lo_count, s.hi_count = g.M, g.N, g.lo_count, g.hi_count
        s.k_blocks, s.k_inner, s.split_k, s.block_n = g.k_blocks, g.k_inner, g.split_k, g.block_n
        esz = g.c.element_size()
        assert g.c.is_cuda and (g.c.dtype == torch.bfloat16) == (g.c_dtype == OUT_BF16)
        s.c = g.c.data_ptr() + esz * g.c_offset
        s.c_dtype, s.act = g.c_dtype, g.act
        s.ldc, s.c_stride_lo, s.c_stride_hi = g.ldc, g.c_stride_lo, g.c_stride_hi
        s.z_out = None if g.z_out is None else g.z_out.data_ptr() + 2 * g.c_offset
        s.aux = None if g.aux is None else g.aux.data_ptr() + 2 * g.c_offset
        s.aux_mode, s.bias_stride_lo = g.aux_mode, g.bias_stride_lo
        if g.bias is not None:
            assert g.bias.dtype == torch.float32 and g.bias.is_cuda
            s.bias = g.bias.data_ptr()
        s.alpha = float(g.alpha)
        _lib.check(self.lib.a8_gemm(C.byref(s), _stream()), "a8_gemm")


    # ------------------------------------------------------------------ ctc
    def ctc_prep(self, targets, pad, eos, target_lengths, input_lengths):
        B, S = targets.shape
        dev = targets.device
        assert targets.dtype == torch.int64 and target_lengths.dtype == torch.int64 and input_lengths.dtype == torch.int64
        ints = torch.empty(B * S + 4 * B, dtype=torch.int32, device=dev)
        flat, row_start, off, tl, il = (ints[:B * S], ints[B * S:B * S + B], ints[B * S + B:B * S + 2 * B],
                                        ints[B * S + 2 * B:B * S + 3 * B], ints[B * S + 3 * B:])
        _lib.check(self.lib.a8_ctc_prep(_ptr(targets), targets.stride(0), targets.stride(1), B, S, pad, eos,
                                        _ptr(target_lengths), _ptr(input_lengths), _ptr(flat), _ptr(row_start),
                                        _ptr(off), _ptr(tl), _ptr(il), _stream()), "a8_ctc_prep")
        return flat, off, tl, il

    def ctc_forward(self, lp, flat, off, tl, il, max_S, blank, mean, zero_inf):
        T, B, V = lp.shape
        assert lp.is_cuda and lp.dtype == torch.float32
        n = self.lib.a8_ctc_scratch_floats(T, B, max_S)
        alpha = torch.empty(n, dtype=torch.float32, device=lp.device)
        beta = torch.empty(n, dtype=torch.float32, device=lp.device)
        nll = torch.empty(B, dtype=torch.float32, device=lp.device)
        loss = torch.empty((), dtype=torch.float32, device=lp.device)
        _lib.check(self.lib.a8_ctc_forward(_ptr(lp), lp.stride(0), lp.stride(1), lp.stride(2), T, B, V, _ptr(flat),
                                           _ptr(off), _ptr(tl), _ptr(il), max_S, blank, int(mean), int(zero_inf),
                                           _ptr(alpha), _ptr(beta), _ptr(nll), _ptr(loss), _stream()),
                   "a8_ctc_forward")
        return loss, nll, alpha, beta

    def ctc_backward(self, lp, flat, off, tl, il, max_S, blank, alpha, beta, nll, grad_out, mean, zero_inf):
        T, B, V = lp.shape
        grad = torch.empty(T, B, V, dtype=torch.float32, device=lp.device)
        go = grad_out.contiguous().float()
        go_stride = 0 if go.numel() == 1 else 1
        _lib.check(self.lib.a8_ctc_backward(_ptr(lp), lp.stride(0), lp.stride(1), lp.stride(2), T, B, V, _ptr(flat),
                                            _ptr(off), _ptr(tl), _ptr(il), max_S, blank, _ptr(alpha), _ptr(beta),
                                            _ptr(nll), _ptr(go), go_stride, int(mean), int(zero_inf), _ptr(grad),
                                            _stream()), "a8_ctc_backward")
        return grad


_BACKEND = None


def backend():
    global _BACKEND
    if _BACKEND is None:
        _BACKEND = CudaBackend()
    return _BACKEND


def set_backend(b):
    """Test hook (tests/emu.py).  The product never calls this."""
    global _BACKEND
    _BACKEND = b


def gemm(spec):
    backend().gemm(spec)
