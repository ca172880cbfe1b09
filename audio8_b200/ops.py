"""Python face of the C ABI: tensor-level wrappers around libaudio8_b200.so.

Every function here takes torch CUDA tensors, checks them, and enqueues hand-written sm_100a kernels on the
current CUDA stream through `ctypes`.  There is no CPU or PyTorch fallback: a missing library or a failed
launch raises.  (`tests/emu.py` contains a pure-PyTorch emulation of the same ABI that CPU tests inject to
check the host-side orchestration without a GPU; the product never imports it.)
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_GELU_DZ, ACT_NONE, AUX_ADD, AUX_MUL, AUX_MUL_GELU_GRAD, AUX_NONE, MAJOR_K, MAJOR_MN, OUT_BF16, OUT_F32,
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
                 aux=None, aux_mode=AUX_NONE, bias=None, bias_stride_lo=0, alpha=1.0, cluster=1, window_k16=None,
                 colsum=None):
        self.a, self.b, self.M, self.N, self.k_blocks = a, b, M, N, k_blocks
        self.k_inner = k_inner if k_inner is not None else k_blocks
        self.c, self.ldc, self.c_dtype, self.c_offset = c, ldc, c_dtype, c_offset
        self.lo_count, self.hi_count, self.split_k, self.block_n = lo_count, hi_count, split_k, block_n
        self.c_stride_lo, self.c_stride_hi = c_stride_lo, c_stride_hi
        self.act, self.z_out, self.aux, self.aux_mode = act, z_out, aux, aux_mode
        self.bias, self.bias_stride_lo, self.alpha = bias, bias_stride_lo, alpha
        self.cluster = cluster  # 2: CTA pair per 256 x BN tile (cta_group::2); block_n 128 / 256 (192 with a K-major B)
        # not None: tap-window kernel (A8_GEMM_TAP_WINDOW), the value = 16-wide k-steps per tap that hold non-zero weights
        self.window_k16 = window_k16
        # fp32 [N] accumulator (zeroed by the caller) that receives the column sums of the stored bf16 output: the bias
        # gradient of the layer whose output gradient this GEMM produces (a8_gemm_t::colsum; AUX_MUL epilogues only)
        self.colsum = colsum
        self.flops = 0  # algorithmic 2*MACs of this launch (set by gemm_specs builders; bench accounting only)


# gradient arena of the data-parallel wrapper (parallel.py): [arena used by forward-time decisions, arena for the
# backward that belongs to the last forward, callback fired when the encoder's backward has been enqueued]
_ARENA = [None, None, None]


def set_grad_arena(arena, encoder_done_cb, keep_for_backward=False):
    if keep_for_backward:  # forward is over: later forwards must not see it, this step's backward still does
        _ARENA[0] = None
        return
    _ARENA[0] = _ARENA[1] = arena
    _ARENA[2] = encoder_done_cb


def grad_arena_for_backward():
    return _ARENA[1]


def grad_arena_serial():
    """0 without an arena, else the serial number of the one this forward runs under.  The captured backward of a CUDA
    graph writes to the addresses it saw at capture time (arena blocks, or the graph's own pool), so graphs.py keeps
    one capture per arena and never replays one whose arena is gone"""
    a = _ARENA[0]
    return 0 if a is None else a.serial


def grad_key(param):
    """arena key of a parameter (its storage address: the aliases a CUDA-graph capture runs on share it); None stays None"""
    return None if param is None else ("p", param.data_ptr())


def grad_arena_take(key, numel, zero=True):
    """the parameter's (or transformer layer's) persistent gradient block of the data-parallel wrapper's arena, zeroed
    unless zero=False; None when no arena is active for this backward or the key has no block (allocate normally)"""
    a = _ARENA[1]
    return a.take(key, numel, zero) if a is not None else None


def encoder_backward_done():
    cb = _ARENA[2]
    if cb is not None and _ARENA[1] is not None:
        cb()


def grad_arena_active():
    return _ARENA[1] is not None


_DYN = [0, 0]  # (rows of the current step's masked-row tensors, their upper bound); set_dynamic_rows()


def set_dynamic_rows(rows, rows_max):
    """The model announces, per step, the masked-row count R = B*Tm and its upper bound for this input shape.
    Tensors whose size is a multiple of R are then allocated with their worst-case size (`bucketed_empty`), so the
    caching allocator sees the same request sizes on every step: no cudaMalloc (tens of ms, device-synchronising)
    shows up in steady state when a step draws a few more masked frames than any step before it."""
    _DYN[0], _DYN[1] = int(rows), int(max(rows, rows_max))


def bucketed_empty(shape, dtype, device, zero=False):
    """torch.empty for tensors whose size follows the per-step masked-row count.  With set_dynamic_rows() in effect
    and an element count divisible by the row count, the worst-case size is allocated and a view of the exact shape
    returned; otherwise the element count is rounded up to 4 size classes per power of two (<= 25 % slack)."""
    if isinstance(shape, int):
        shape = (shape,)
    n = 1
    for d in shape:
        n *= int(d)
    R, Rmax = _DYN
    if R > 0 and n > 0 and n % R == 0:
        cap = n // R * Rmax
    else:
        if n * torch.empty((), dtype=dtype).element_size() < 65536:
            return torch.zeros(shape, dtype=dtype, device=device) if zero else torch.empty(shape, dtype=dtype, device=device)
        g = 1 << max(n.bit_length() - 3, 0)
        cap = (n + g - 1) // g * g
    buf = torch.empty(cap, dtype=dtype, device=device)
    out = buf[:n]
    if zero:
        out.zero_()
    return out.view(shape)


def _stream():
    # raw cudaStream_t of torch's current stream on the current device (cheaper than current_stream().cuda_stream)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def _ptr(t, offset_elems=0):
    """device address as a plain int (ctypes converts it for c_void_p parameters); None stays NULL"""
    if t is None:
        return None
    return t.data_ptr() + offset_elems * t.element_size()


class _FrozenGemm:
    """prebuilt a8_gemm_t + the recipe to refresh its pointer fields from a call's tensor arguments"""

    __slots__ = ("struct", "ref", "slots", "flops")


class CudaBackend:
    """Calls into libaudio8_b200.so.  Constructed lazily; raises if the library is absent."""

    name = "cuda"

    def __init__(self):
        self.lib = _lib.load()
        self.profiler = None  # bench.py installs a CUDA-event recorder around GEMM launches

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
        if not isinstance(g, GemmSpec):
            return self._gemm_bound(g)
        self._gemm_spec(g)

    def _gemm_bound(self, b):
        fz = b.cache.get(b.key)
        tens = b.tensors()
        if fz is None:
            spec = b.spec()
            st = self._fill(spec)
            fz = _FrozenGemm()
            fz.struct, fz.flops = st, spec.flops
            fz.ref = C.byref(st)
            # pointer field -> (index of the argument tensor that owns it, byte offset from that tensor's data_ptr)
            fields = [("a", spec.a.t, st.a.ptr), ("b", spec.b.t, st.b.ptr), ("c", spec.c, st.c), ("z", spec.z_out, st.z_out),
                      ("x", spec.aux, st.aux), ("s", spec.bias, st.bias), ("q", spec.colsum, st.colsum)]
            fz.slots = []
            for name, t, val in fields:
                if t is None:
                    continue
                owner = [i for i, u in enumerate(tens) if u is t]
                if not owner:  # the builder derived a new tensor: cannot be frozen
                    fz = None
                    break
                fz.slots.append((name, owner[0], val - t.data_ptr()))
            if fz is None:
                return self._gemm_spec(spec)
            b.cache[b.key] = fz
        st = fz.struct
        for name, i, delta in fz.slots:
            p = tens[i].data_ptr() + delta
            if name == "a":
                st.a.ptr = p
            elif name == "b":
                st.b.ptr = p
            elif name == "c":
                st.c = p
            elif name == "z":
                st.z_out = p
            elif name == "x":
                st.aux = p
            elif name == "q":
                st.colsum = p
            else:
                st.bias = p
        if self.profiler is not None:
            self.profiler.begin("gemm", fz.flops)
        _lib.check(self.lib.a8_gemm(fz.ref, _stream()), "a8_gemm")
        if self.profiler is not None:
            self.profiler.end()

    def _gemm_spec(self, g):
        s = self._fill(g)
        if self.profiler is not None:
            self.profiler.begin("gemm", g.flops)
        _lib.check(self.lib.a8_gemm(C.byref(s), _stream()), "a8_gemm")
        if self.profiler is not None:
            self.profiler.end()

    def gemm_group(self, specs):
        """one persistent launch per <= 48 problems (a8_gemm_group): same majors / tiling / epilogue kind, own operands.
        The prepared launch (tensor maps + tile table, ~100 us of host time) is cached on the operand addresses: the
        caching allocator hands the same blocks back step after step, and CUDA-graph capture sees it once anyway."""
        cache = self.__dict__.setdefault("_group_cache", {})
        for lo in range(0, len(specs), 48):
            part = [b.spec() if hasattr(b, "spec") else b for b in specs[lo:lo + 48]]
            key = tuple((g.a.t.data_ptr(), g.b.t.data_ptr(), g.c.data_ptr(), g.M, g.N, g.k_blocks, g.c_dtype, g.split_k)
                        for g in part)
            ent = cache.get(key)
            if ent is None:
                arr = (_lib.Gemm * len(part))()
                for i, g in enumerate(part):
                    arr[i] = self._fill(g)
                blob = C.create_string_buffer(self.lib.a8_gemm_group_blob_bytes())
                _lib.check(self.lib.a8_gemm_group_prepare(arr, len(part), blob), "a8_gemm_group_prepare")
                if len(cache) >= 8:
                    cache.pop(next(iter(cache)))
                ent = cache[key] = (blob, sum(g.flops for g in part))
            if self.profiler is not None:
                self.profiler.begin("gemm", ent[1])
            _lib.check(self.lib.a8_gemm_group_launch(ent[0], _stream()), "a8_gemm_group_launch")
            if self.profiler is not None:
                self.profiler.end()

    def _fill(self, g):
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
        if g.colsum is not None:
            assert g.colsum.dtype == torch.float32 and g.colsum.is_cuda and g.colsum.numel() >= g.N and g.aux_mode == AUX_MUL
            s.colsum = g.colsum.data_ptr()
        s.reserved = 2 if g.cluster == 2 else 0
        if g.window_k16 is not None:
            s.reserved = 3 | (int(g.window_k16) << 8)
        return s


    # ------------------------------------------------------------------ ctc
    def ctc_prep(self, targets, pad, eos, target_lengths, input_lengths):
        B, S = targets.shape
        dev = targets.device
        assert targets.dtype == torch.int64 and target_lengths.dtype == torch.int64 and input_lengths.dtype == torch.int64
        ints = torch.empty(B * S + 4 * B + 1, dtype=torch.int32, device=dev)
        flat, row_start, off, tl, il = (ints[:B * S], ints[B * S:B * S + B], ints[B * S + B:B * S + 2 * B],
                                        ints[B * S + 2 * B:B * S + 3 * B], ints[B * S + 3 * B:])  # il: B lengths + ticket
        _lib.check(self.lib.a8_ctc_prep(_ptr(targets), targets.stride(0), targets.stride(1), B, S, pad, eos,
                                        _ptr(target_lengths), _ptr(input_lengths), _ptr(flat), _ptr(row_start),
                                        _ptr(off), _ptr(tl), _ptr(il), _stream()), "a8_ctc_prep")
        return flat, off, tl, il

    def ctc_forward(self, x, flat, off, tl, il, max_S, blank, mean, zero_inf, from_logits=False):
        """x: fp32 [T,B,V] view (any strides) of log-probs, or of the classifier's logits when from_logits"""
        T, B, V = x.shape
        assert x.is_cuda and x.dtype == torch.float32 and il.numel() == B + 1
        n = self.lib.a8_ctc_scratch_floats(T, B, max_S)
        ab = torch.empty(2, n, dtype=torch.float32, device=x.device)  # alpha | beta~ scratch
        out = torch.empty(B + 1, dtype=torch.float32, device=x.device)
        nll, loss = out[:B], out[B]
        _lib.check(self.lib.a8_ctc_forward(_ptr(x), x.stride(0), x.stride(1), x.stride(2), T, B, V, int(from_logits),
                                           _ptr(flat), _ptr(off), _ptr(tl), _ptr(il), max_S, blank, int(mean),
                                           int(zero_inf), _ptr(ab[0]), _ptr(ab[1]), _ptr(nll), _ptr(loss), _stream()),
                   "a8_ctc_forward")
        return loss, nll, ab

    def ctc_backward(self, x, flat, off, tl, il, max_S, blank, alpha, nll, grad_out, mean, zero_inf, from_logits=False,
                     batch_major=False):
        """-> fp32 gradient w.r.t. x, laid out [T,B,V] contiguous, or [B,T,V] contiguous (returned as its [T,B,V] view)
        when batch_major: the layout of the logits tensor it flows back into"""
        T, B, V = x.shape
        if batch_major:
            grad = torch.empty(B, T, V, dtype=torch.float32, device=x.device).transpose(0, 1)
        else:
            grad = torch.empty(T, B, V, dtype=torch.float32, device=x.device)
        go = grad_out.contiguous().float()
        go_stride = 0 if go.numel() == 1 else 1
        _lib.check(self.lib.a8_ctc_backward(_ptr(x), x.stride(0), x.stride(1), x.stride(2), T, B, V, int(from_logits),
                                            _ptr(flat), _ptr(off), _ptr(tl), _ptr(il), max_S, blank, _ptr(alpha[0]),
                                            _ptr(alpha[1]), _ptr(nll), _ptr(go), go_stride, int(mean), int(zero_inf), _ptr(grad),
                                            grad.stride(0), grad.stride(1), _stream()), "a8_ctc_backward")
        return grad

    def ctc_greedy(self, lp, in_len, blank):
        """lp fp32 [B,T,V] (any strides), in_len int32 [B] or None -> (ids int32 [B,T] padded with -1, lengths int32 [B])"""
        B, T, V = lp.shape
        assert lp.is_cuda and lp.dtype == torch.float32 and (in_len is None or in_len.dtype == torch.int32)
        out = torch.empty(B, T, dtype=torch.int32, device=lp.device)
        out_len = torch.empty(B, dtype=torch.int32, device=lp.device)
        _lib.check(self.lib.a8_ctc_greedy(_ptr(lp), lp.stride(0), lp.stride(1), lp.stride(2), B, T, V, _ptr(in_len), blank,
                                          _ptr(out), _ptr(out_len), _stream()), "a8_ctc_greedy")
        return out, out_len

    # ------------------------------------------------------------------ dropout seeds
    def set_seed_source(self, t):
        """t: int64 CUDA tensor with one element (or None): added to the seed argument of every later dropout-capable
        launch, read on the device at run time (CUDA-graph replays see the refreshed value)"""
        assert t is None or (t.is_cuda and t.dtype == torch.int64 and t.numel() >= 1)
        self.lib.a8_set_seed_source(_ptr(t))

    # ------------------------------------------------------------------ row kernels
    def layernorm_fwd(self, x, gamma, beta, eps, h=None, p_h=0.0, seed_h=0, want_f32=False, p_y=0.0, seed_y=0):
        """returns y (bf16), y_f32 or None, s (bf16: x + drop(h), or x itself when h is None), mean, rstd"""
        C = x.shape[-1]
        R = x.numel() // C
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.is_cuda
        y = torch.empty_like(x)
        yf = torch.empty(x.shape, dtype=torch.float32, device=x.device) if want_f32 else None
        s = torch.empty_like(x) if h is not None else x
        mean = torch.empty(R, dtype=torch.float32, device=x.device)
        rstd = torch.empty(R, dtype=torch.float32, device=x.device)
        _lib.check(self.lib.a8_layernorm_fwd(_ptr(x), _ptr(h), p_h, seed_h, _ptr(s) if h is not None else None,
                                             _ptr(gamma), _ptr(beta), eps, _ptr(y), _ptr(yf), p_y, seed_y, _ptr(mean),
                                             _ptr(rstd), R, C, _stream()), "a8_layernorm_fwd")
        return y, yf, s, mean, rstd

    def layernorm_bwd(self, dy, s, mean, rstd, gamma, dy_f32=None, p_y=0.0, seed_y=0, want_dh=False, p_h=0.0,
                      seed_h=0, want_dbias=False, acc=None, dg_out=None, db_out=None):
        """returns ds, dh (or None), dgamma, dbeta, dbias (or None); acc: optional ZEROED fp32 [3,C] accumulator;
        dg_out / db_out: optional ZEROED fp32 [C] destinations for dgamma / dbeta (gradient-arena blocks)"""
        C = s.shape[-1]
        R = s.numel() // C
        assert dy.dtype == torch.bfloat16 and dy.is_contiguous() and s.is_contiguous()
        ds = torch.empty_like(s)
        dh = torch.empty_like(s) if want_dh else None
        if dg_out is not None and db_out is not None and not want_dbias:
            a0, a1, a2 = dg_out, db_out, None
        else:
            if acc is None:
                acc = torch.zeros(3, C, dtype=torch.float32, device=s.device)
            a0, a1, a2 = acc[0], acc[1], acc[2]
        _lib.check(self.lib.a8_layernorm_bwd(_ptr(dy), _ptr(dy_f32), p_y, seed_y, _ptr(s), _ptr(mean), _ptr(rstd),
                                             _ptr(gamma), _ptr(ds), _ptr(dh), p_h, seed_h, _ptr(a0), _ptr(a1),
                                             _ptr(a2) if want_dbias else None, R, C, _stream()),
                   "a8_layernorm_bwd")
        return ds, dh, a0, a1, (a2 if want_dbias else None)

    # ------------------------------------------------------------------ fused attention
    def attn_fwd(self, qkv, H, scale, key_keep=None, pdrop=0.0, seed=0):
        """qkv bf16 [B,T,3D] -> (ctx bf16 [B,T,D], lse fp32 [B,H,T]); scores / probabilities stay on chip"""
        B, T, D3 = qkv.shape
        D = D3 // 3
        assert qkv.is_cuda and qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and D == H * 64
        ctx = torch.empty(B, T, D, dtype=torch.bfloat16, device=qkv.device)
        lse = torch.empty(B, H, T, dtype=torch.float32, device=qkv.device)
        _lib.check(self.lib.a8_attn_fwd(_ptr(qkv), _ptr(key_keep), _ptr(ctx), _ptr(lse), B, H, T, scale, pdrop, seed,
                                        _stream()), "a8_attn_fwd")
        return ctx, lse

    def attn_bwd(self, qkv, ctx, dctx, lse, H, scale, key_keep=None, pdrop=0.0, seed=0, dbias=None):
        """-> dqkv bf16 [B,T,3D] (dQ | dK | dV); dbias (fp32 [3D], zeroed by the caller) += its column sums"""
        B, T, D3 = qkv.shape
        assert dctx.dtype == torch.bfloat16 and dctx.is_contiguous() and ctx.is_contiguous() and dctx.shape == ctx.shape
        assert dbias is None or (dbias.dtype == torch.float32 and dbias.is_cuda and dbias.numel() == D3 and dbias.is_contiguous())
        dqkv = torch.empty_like(qkv)
        delta = torch.empty(B, H, T, dtype=torch.float32, device=qkv.device)
        _lib.check(self.lib.a8_attn_bwd(_ptr(qkv), _ptr(key_keep), _ptr(ctx), _ptr(dctx), _ptr(lse), _ptr(delta),
                                        _ptr(dqkv), _ptr(dbias), B, H, T, scale, pdrop, seed, _stream()), "a8_attn_bwd")
        return dqkv

    def attn_dropmask(self, B, H, T, pdrop, seed, device):
        out = torch.empty(B, H, T, T, dtype=torch.uint8, device=device)
        _lib.check(self.lib.a8_attn_dropmask(_ptr(out), B, H, T, pdrop, seed, _stream()), "a8_attn_dropmask")
        return out

    def colsum(self, x, out=None):
        """out: optional ZEROED fp32 [C] accumulator"""
        C = x.shape[-1]
        R = x.numel() // C
        assert x.dtype == torch.bfloat16 and x.is_contiguous()
        if out is None:
            out = torch.zeros(C, dtype=torch.float32, device=x.device)
        _lib.check(self.lib.a8_colsum(_ptr(x), C, R, C, _ptr(out), _stream()), "a8_colsum")
        return out

    def dropout(self, x, p, seed):
        assert x.is_contiguous()
        out = bucketed_empty(x.shape, x.dtype, x.device)
        _lib.check(self.lib.a8_dropout(_ptr(x), _ptr(out), self._dt(x.dtype), x.numel(), p, seed, _stream()),
                   "a8_dropout")
        return out

    def gelu_bwd(self, dy, z):
        assert dy.dtype == torch.bfloat16 and z.dtype == torch.bfloat16 and dy.is_contiguous() and z.is_contiguous()
        dz = torch.empty_like(z)
        _lib.check(self.lib.a8_gelu_bwd(_ptr(dy), _ptr(z), _ptr(dz), z.numel(), _stream()), "a8_gelu_bwd")
        return dz

    def mul_dgelu(self, dy, g):
        """GELU backward with the stored derivative: dy (bf16) * g (fp16: the gelu'(z) a forward GEMM wrote with ACT_GELU_DZ)"""
        assert dy.dtype == torch.bfloat16 and g.dtype == torch.float16 and dy.is_contiguous() and g.is_contiguous()
        out = torch.empty_like(dy)
        _lib.check(self.lib.a8_mul_dgelu(_ptr(dy), _ptr(g), _ptr(out), dy.numel(), _stream()), "a8_mul_dgelu")
        return out

    def log_softmax_fwd(self, x):
        V = x.shape[-1]
        assert x.dtype == torch.float32 and x.is_contiguous()
        y = torch.empty_like(x)
        _lib.check(self.lib.a8_log_softmax_fwd(_ptr(x), _ptr(y), x.numel() // V, V, _stream()), "a8_log_softmax_fwd")
        return y

    def log_softmax_bwd(self, dy, y):
        """dy: fp32 gradient w.r.t. y [B,T,V], any strides (e.g. the transposed [T,B,V] tensor CTC returns)"""
        B, T, V = y.shape
        assert dy.shape == y.shape and dy.dtype == torch.float32 and y.is_contiguous()
        dx = torch.empty(y.shape, dtype=torch.bfloat16, device=y.device)
        _lib.check(self.lib.a8_log_softmax_bwd(_ptr(dy), dy.stride(0), dy.stride(1), dy.stride(2), T, _ptr(y), _ptr(dx),
                                               B * T, V, _stream()), "a8_log_softmax_bwd")
        return dx

    # ------------------------------------------------------------------ conv layer 0
    def conv0_stats(self, x, w, k, stride, eps):
        B, L = x.shape
        C = w.shape[0]
        assert x.dtype == torch.float32 and x.is_contiguous() and w.is_contiguous()
        mom = torch.empty(65 * B, dtype=torch.float64, device=x.device)
        mean = torch.empty(B, C, dtype=torch.float32, device=x.device)
        rstd = torch.empty(B, C, dtype=torch.float32, device=x.device)
        _lib.check(self.lib.a8_conv0_stats(_ptr(x), B, L, _ptr(w), C, k, stride, eps, _ptr(mom), _ptr(mean), _ptr(rstd),
                                           _stream()), "a8_conv0_stats")
        return mean, rstd, mom

    def conv0_fwd(self, x, w, gamma, beta, mean, rstd, k, stride):
        B, L = x.shape
        C = w.shape[0]
        L0 = (L - k) // stride + 1
        y = torch.empty(B, L0, C, dtype=torch.bfloat16, device=x.device)
        _lib.check(self.lib.a8_conv0_fwd(_ptr(x), B, L, _ptr(w), _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(rstd), C, k,
                                         stride, _ptr(y), _stream()), "a8_conv0_fwd")
        return y

    def conv0_bwd(self, x, w, gamma, beta, mean, rstd, mom, k, stride, da, out=None):
        """out: optional (dw [C,k], dgamma [C], dbeta [C]) fp32 destinations (fully overwritten)"""
        B, L = x.shape
        C = w.shape[0]
        assert da.dtype == torch.bfloat16 and da.is_contiguous() and mom.dtype == torch.float64
        acc = torch.empty(B * C * 12, dtype=torch.float32, device=x.device)
        if out is None:
            buf = torch.empty(C * k + 2 * C, dtype=torch.float32, device=x.device)
            out = (buf[:C * k].view(C, k), buf[C * k:C * k + C], buf[C * k + C:])
        dw, dg, db = out
        _lib.check(self.lib.a8_conv0_bwd(_ptr(x), B, L, _ptr(w), _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(rstd),
                                         _ptr(mom), C, k, stride, _ptr(da), _ptr(acc), _ptr(dw), _ptr(dg), _ptr(db),
                                         _stream()), "a8_conv0_bwd")
        return dw, dg, db

    # ------------------------------------------------------------------ masks / indices / casts
    @staticmethod
    def _dt(t):
        return {torch.float32: 0, torch.bfloat16: 1}[t]

    def rows_gather(self, src, idx, out_dtype):
        C = src.shape[-1]
        assert src.is_contiguous() and idx.dtype == torch.int32
        out = bucketed_empty((idx.numel(), C), out_dtype, src.device)
        _lib.check(self.lib.a8_rows_copy(_ptr(src), self._dt(src.dtype), _ptr(out), self._dt(out_dtype), _ptr(idx),
                                         idx.numel(), C, 0, _stream()), "a8_rows_copy")
        return out

    def rows_scatter(self, src, idx, n_rows, out_dtype):
        """zero-filled [n_rows, C] with out[idx[i]] = src[i]"""
        C = src.shape[-1]
        assert src.is_contiguous() and idx.dtype == torch.int32
        out = torch.zeros(n_rows, C, dtype=out_dtype, device=src.device)
        _lib.check(self.lib.a8_rows_copy(_ptr(src), self._dt(src.dtype), _ptr(out), self._dt(out_dtype), _ptr(idx),
                                         idx.numel(), C, 1, _stream()), "a8_rows_copy")
        return out

    def rows_set(self, x, idx, vec):
        """in place: x[idx[i], :] = vec"""
        C = x.shape[-1]
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and vec.dtype == torch.float32
        _lib.check(self.lib.a8_rows_set(_ptr(x), _ptr(idx), idx.numel(), C, _ptr(vec), _stream()), "a8_rows_set")

    def rows_set_bwd(self, dx, idx, out=None):
        """in place: zero dx[idx[i], :]; returns the column sum of the rows it zeroed (out: optional ZEROED fp32 [C])"""
        C = dx.shape[-1]
        dvec = out if out is not None else torch.zeros(C, dtype=torch.float32, device=dx.device)
        _lib.check(self.lib.a8_rows_set_bwd(_ptr(dx), _ptr(idx), idx.numel(), C, _ptr(dvec), _stream()),
                   "a8_rows_set_bwd")
        return dvec

    def mask_apply(self, x, row_keep=None, chan_zero=None):
        """in place on x bf16 [B,T,C]: zero rows with row_keep == 0 and channels with chan_zero != 0"""
        B, T, C = x.shape
        _lib.check(self.lib.a8_mask_apply(_ptr(x), _ptr(row_keep), _ptr(chan_zero), B, T, C, _stream()), "a8_mask_apply")

    def cast(self, x, dtype, out=None):
        assert x.is_contiguous()
        if out is None:
            out = bucketed_empty(x.shape, dtype, x.device)
        _lib.check(self.lib.a8_cast(_ptr(x), self._dt(x.dtype), _ptr(out), self._dt(dtype), x.numel(), _stream()),
                   "a8_cast")
        return out

    def split3(self, x, b_side, out=None):
        R, C = x.shape
        assert x.dtype == torch.float32 and x.is_contiguous()
        if out is None:
            out = bucketed_empty((R, 3 * C), torch.bfloat16, x.device)
        _lib.check(self.lib.a8_split3(_ptr(x), _ptr(out), R, C, int(b_side), _stream()), "a8_split3")
        return out

    # ------------------------------------------------------------------ parameter re-layout
    def cast_multi(self, pairs, cache):
        """pairs: [(src fp32, dst bf16|fp32)] with equal numel; the device table is cached in `cache` while the
        pointers stay the same (parameters are updated in place by optimizers)"""
        sig = tuple((s.data_ptr(), d.data_ptr()) for s, d in pairs)
        if cache.get("sig") != sig:
            rows = [[s.data_ptr(), d.data_ptr(), s.numel(), 1 if d.dtype == torch.float32 else 0] for s, d in pairs]
            for (s, d) in pairs:
                assert s.dtype == torch.float32 and s.is_contiguous() and d.is_contiguous() and s.numel() == d.numel()
            cache["table"] = torch.tensor(rows, dtype=torch.int64).to(pairs[0][0].device)
            cache["sig"] = sig
        _lib.check(self.lib.a8_cast_multi(_ptr(cache["table"]), len(pairs), _stream()), "a8_cast_multi")

    def conv_pack(self, w, s, want_t, out=None):
        """[Cout,Cin,k] fp32 -> (wk bf16 [Cout,k*Cin], [wt_0, wt_1] bf16 [Cin, ntaps_p*Cout] or None); out: a previous
        result to overwrite (persistent operand buffers)"""
        Cout, Cin, k = w.shape
        assert w.dtype == torch.float32 and w.is_contiguous() and s == 2
        if out is not None:
            wk, wts = out
        else:
            wk = torch.empty(Cout, k * Cin, dtype=torch.bfloat16, device=w.device)
            wts = None
            if want_t:
                wts = [torch.empty(Cin, ((k - p + s - 1) // s) * Cout, dtype=torch.bfloat16, device=w.device) for p in range(s)]
        _lib.check(self.lib.a8_conv_pack(_ptr(w), Cout, Cin, k, s, _ptr(wk), _ptr(wts[0]) if wts else None,
                                         _ptr(wts[1]) if wts else None, _stream()), "a8_conv_pack")
        return wk, wts

    def conv_unpack(self, dwk, Cin, k, out=None):
        Cout = dwk.shape[0]
        dw = out if out is not None else torch.empty(Cout, Cin, k, dtype=torch.float32, device=dwk.device)
        _lib.check(self.lib.a8_conv_unpack(_ptr(dwk), Cout, Cin, k, _ptr(dw), _stream()), "a8_conv_unpack")
        return dw

    def posconv_pack(self, g, v, want_t, out=None):
        """weight-normed pos-conv weight -> (wp, wpt or None, norm2); out: a previous result to overwrite"""
        D, cg, k = v.shape
        assert g.numel() == k and v.is_contiguous() and g.is_contiguous()
        if out is not None:
            wp, wpt, norm2 = out  # norm2 is a view of the buffer that also holds the reduction scratch
        else:
            buf = torch.empty(k + self.lib.a8_posconv_norm_scratch_floats(D, cg, k), dtype=torch.float32, device=v.device)
            norm2 = buf[:k]  # followed by the ordered-partial-sum scratch of the deterministic reduction
            wp = torch.empty(D, k * 64, dtype=torch.bfloat16, device=v.device)
            wpt = torch.empty(D, k * 64, dtype=torch.bfloat16, device=v.device) if want_t else None
        _lib.check(self.lib.a8_posconv_pack(_ptr(g), _ptr(v), D, cg, k, _ptr(norm2), _ptr(wp), _ptr(wpt), _stream()),
                   "a8_posconv_pack")
        return wp, wpt, norm2

    def posconv_wn_bwd(self, dwp, g, v, norm2, out=None):
        """out: optional (dv like v, dg like g) fp32 destinations (fully overwritten)"""
        D, cg, k = v.shape
        t = torch.empty(k, dtype=torch.float32, device=v.device)
        dv, dg = out if out is not None else (torch.empty_like(v), torch.empty_like(g))
        _lib.check(self.lib.a8_posconv_wn_bwd(_ptr(dwp), _ptr(g), _ptr(v), _ptr(norm2), D, cg, k, _ptr(t), _ptr(dv),
                                              _ptr(dg), _stream()), "a8_posconv_wn_bwd")
        return dv, dg

    # ------------------------------------------------------------------ quantizer / contrastive
    def vq_fwd(self, z, noise, tau, vars2d, G, n_valid=None):
        R = z.shape[0]
        V = z.shape[1] // G
        vd = vars2d.shape[1]
        dev = z.device
        q = bucketed_empty((R, G * vd), torch.float32, dev)
        qb = bucketed_empty((R, G * vd), torch.bfloat16, dev)
        kidx = torch.empty(R * G, dtype=torch.int32, device=dev)
        avg = torch.empty(V, dtype=torch.float32, device=dev)
        ppl = torch.empty((), dtype=torch.float32, device=dev)
        _lib.check(self.lib.a8_vq_fwd(_ptr(z), _ptr(noise), tau, _ptr(vars2d), R, G, V, vd, _ptr(n_valid), _ptr(q), _ptr(qb), _ptr(kidx),
                                      _ptr(avg), _ptr(ppl), _stream()), "a8_vq_fwd")
        return q, qb, kidx, avg, ppl

    def vq_bwd(self, z, noise, tau, G, vd, a_dot, dq, kidx, avg, ppl, dppl, dvars_out=None, n_valid=None):
        """dvars_out: optional ZEROED fp32 [G*V, vd] destination of the codebook gradient"""
        R = z.shape[0]
        V = z.shape[1] // G
        dz = bucketed_empty((R, G * V), torch.bfloat16, z.device)
        dvars = dvars_out if dvars_out is not None else torch.zeros(G * V, vd, dtype=torch.float32, device=z.device)
        _lib.check(self.lib.a8_vq_bwd(_ptr(z), _ptr(noise), tau, R, G, V, vd, _ptr(n_valid), _ptr(a_dot), _ptr(dq), _ptr(kidx), _ptr(avg),
                                      _ptr(ppl), _ptr(dppl), _ptr(dz), _ptr(dvars), _stream()), "a8_vq_bwd")
        return dz, dvars

    def contrastive_fwd(self, x, y, idx, ppl, n_vars, xe_w, div_w, n_valid=None):
        R, Cc = x.shape
        K = idx.numel() // R
        dev = x.device
        assert x.dtype == torch.float32 and y.dtype == torch.float32 and idx.dtype == torch.int32
        xn = torch.empty(2 * R, dtype=torch.float32, device=dev)
        cp = bucketed_empty((2, R, K + 1), torch.float32, dev)
        rl = torch.empty(R + 2, dtype=torch.float32, device=dev)
        _lib.check(self.lib.a8_contrastive_fwd(_ptr(x), _ptr(y), _ptr(idx), R, Cc, K, _ptr(n_valid), _ptr(ppl), n_vars, xe_w, div_w,
                                               _ptr(xn), _ptr(xn, R), _ptr(cp[0]), _ptr(cp[1]), _ptr(rl), _ptr(rl, R),
                                               _ptr(rl, R + 1), _stream()), "a8_contrastive_fwd")
        return rl[R + 1], rl[R], (xn, cp)

    def contrastive_bwd(self, x, y, idx, saved, dce, n_valid=None):
        R, Cc = x.shape
        K = idx.numel() // R
        xn, cp = saved
        dx = bucketed_empty(x.shape, x.dtype, x.device)
        dy = bucketed_empty(y.shape, y.dtype, y.device)
        _lib.check(self.lib.a8_contrastive_bwd(_ptr(x), _ptr(y), _ptr(idx), R, Cc, K, _ptr(n_valid), _ptr(xn), _ptr(xn, R), _ptr(cp[0]),
                                               _ptr(cp[1]), _ptr(dce), _ptr(dx), _ptr(dy), _stream()),
                   "a8_contrastive_bwd")
        return dx, dy

    # ------------------------------------------------------------------ device-side draws (csrc/draws.cu)
    def span_mask_draw(self, seed, seed_dev, B, T, p_start, mask_length, R_max, device):
        """-> (rows int32 [R_max + 1]: masked flat rows, -1 padding, count last; mask uint8 [B, T])"""
        assert seed_dev is None or (seed_dev.is_cuda and seed_dev.dtype == torch.int64)
        rows = torch.empty(R_max + 1, dtype=torch.int32, device=device)
        mask = torch.empty(B, T, dtype=torch.uint8, device=device)
        _lib.check(self.lib.a8_span_mask_draw(seed, _ptr(seed_dev), B, T, float(p_start), mask_length, R_max, _ptr(rows),
                                              _ptr(mask), _stream()), "a8_span_mask_draw")
        return rows, mask

    def negatives_draw(self, seed, seed_dev, rows, B, K):
        """rows: the padded list of span_mask_draw (its last element is the valid count) -> int32 [R_max * K]"""
        assert rows.is_cuda and rows.dtype == torch.int32 and rows.is_contiguous()
        assert seed_dev is None or (seed_dev.is_cuda and seed_dev.dtype == torch.int64)
        R_max = rows.numel() - 1
        out = torch.empty(R_max * K, dtype=torch.int32, device=rows.device)
        _lib.check(self.lib.a8_negatives_draw(seed, _ptr(seed_dev), _ptr(rows, R_max), B, K, R_max, _ptr(out), _stream()),
                   "a8_negatives_draw")
        return out

    # ------------------------------------------------------------------ optimizer side (csrc/optim.cu)
    def optim_grad_sqnorm(self, table, chunk_tensor, chunk_off, chunk, partials):
        _lib.check(self.lib.a8_optim_grad_sqnorm(_ptr(table), _ptr(chunk_tensor), _ptr(chunk_off), chunk_tensor.numel(),
                                                 chunk, _ptr(partials), _stream()), "a8_optim_grad_sqnorm")

    def optim_adamw(self, table, chunk_tensor, chunk_off, chunk, partials, max_norm, grad_scale, lr, beta1, beta2, eps,
                    weight_decay, bc1, bc2_sqrt, scale_grads_only, total_norm_out):
        _lib.check(self.lib.a8_optim_adamw(_ptr(table), _ptr(chunk_tensor), _ptr(chunk_off), chunk_tensor.numel(), chunk,
                                           _ptr(partials), max_norm, grad_scale, lr, beta1, beta2, eps, weight_decay,
                                           bc1, bc2_sqrt, int(scale_grads_only), _ptr(total_norm_out), _stream()),
                   "a8_optim_adamw")


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
