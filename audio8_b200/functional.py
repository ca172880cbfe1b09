"""autograd Functions of the wav2vec2 training path.

Each Function's forward and backward is a hand-ordered sequence of launches into libaudio8_b200.so (through
`ops.backend()`); PyTorch only owns the tensors, the streams and the outer autograd graph between these
few coarse nodes.  Activations are bf16 channels-last, parameters stay fp32 (their bf16 copies are made per
call), parameter gradients are returned in fp32.

Reference call sites are cited per Function (paths under /root/reference/audio8).
"""
import math

import torch

from . import gemm_specs as G
from . import ops
from .ops import ACT_GELU, ACT_GELU_DZ, ACT_NONE, AUX_ADD, AUX_MUL, AUX_NONE, OUT_BF16, OUT_F32

BF16, F32, F16 = torch.bfloat16, torch.float32, torch.float16
_site = [0]


def next_seed():
    """per-call-site dropout seed (a host constant, baked into CUDA-graph captures); the per-step randomness comes
    from `step_seed()`, a device word that kernels add to it at run time"""
    _site[0] = (_site[0] + 0x9E3779B97F4A7C15) & 0x3FFFFFFFFFFFFFFF
    return _site[0]


def step_seed(like):
    """one int64 on the device from torch's CUDA generator: reproducible under torch.manual_seed() and refreshed on
    every replay when the surrounding code is captured in a CUDA graph (the generator is graph-safe)"""
    return torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=like.device)


def _be():
    return ops.backend()


def _take_saved(ctx):
    """Activations a Function kept for backward, released as backward starts (what save_for_backward does for built-in
    nodes): a loss tensor that outlives its step must not pin the step's activations.  A second backward through the
    same graph (retain_graph) is therefore not supported, as with freed PyTorch buffers."""
    saved = ctx.saved
    if saved is None:
        raise RuntimeError("audio8_b200: backward through this graph a second time: its activations were released "
                           "after the first backward (retain_graph is not supported)")
    ctx.saved = None
    return saved


def _empty(shape, dtype, like, dynamic=False):
    """dynamic=True: the size follows the per-step masked-row count -> bucketed allocation (ops.bucketed_empty)"""
    if dynamic:
        return ops.bucketed_empty(tuple(shape), dtype, like.device)
    return torch.empty(shape, dtype=dtype, device=like.device)


def _zeros(shape, dtype, like, dynamic=False):
    if dynamic:
        return ops.bucketed_empty(tuple(shape), dtype, like.device, zero=True)
    return torch.zeros(shape, dtype=dtype, device=like.device)


def _grad_zeros(key, shape, like):
    """zeroed fp32 buffer for a parameter gradient: the parameter's own block of the data-parallel wrapper's gradient
    arena when one is active for this backward (parallel.py: the gradient is then reduced in place), else a fresh one"""
    n = 1
    for d in shape:
        n *= int(d)
    buf = ops.grad_arena_take(key, n) if key is not None else None
    if buf is None:
        return torch.zeros(shape, dtype=F32, device=like.device)
    return buf.view(shape)


def _grad_empty(key, shape, like):
    """same, for gradients that are fully overwritten by their kernel (no zero fill)"""
    n = 1
    for d in shape:
        n *= int(d)
    buf = ops.grad_arena_take(key, n, zero=False) if key is not None else None
    if buf is None:
        return torch.empty(shape, dtype=F32, device=like.device)
    return buf.view(shape)


def operands(cache, key, params, build):
    """GEMM operands derived from fp32 parameters (bf16 copies, packed conv / pos-conv layouts, bf16x3 splits): rebuilt
    only when a parameter changed (storage address or autograd version counter, which every in-place update bumps),
    and then written into the SAME buffers, so CUDA-graph replays that read them stay valid.  `cache` is a dict owned by
    the module (None: no caching); `build(prev)` launches the re-layout kernels into `prev` when it is not None.
    Modules refresh their operands eagerly at the start of forward(), before any graph segment replays; the calls
    inside the autograd Functions then hit."""
    if cache is None:
        return build(None)
    ver = tuple((p.data_ptr(), p._version) for p in params)
    ent = cache.get(key)
    if ent is not None and ent[0] == ver:
        return ent[1]
    bufs = build(ent[1] if ent is not None else None)
    cache[key] = (ver, bufs)
    return bufs


def linear_operands(cache, weight, bias):
    """(bf16 weight [N8, K], fp32 bias [N8] or None): N padded to a multiple of 8 with zero rows when needed"""
    def build(prev):
        N = weight.shape[0]
        N8 = (N + 7) // 8 * 8
        w32 = weight.detach().contiguous().float()
        if N8 != N:
            w32 = _pad_rows8(w32, N8)
        wb = _be().cast(w32, BF16, out=prev[0] if prev is not None else None)
        bv = None
        if bias is not None:
            bv = bias.detach().float()
            if N8 != N:
                bv = _pad_rows8(bv, N8)
                if prev is not None:
                    prev[1].copy_(bv)
                    bv = prev[1]
        return wb, bv
    if weight.dtype == BF16:
        return weight.detach().contiguous(), (bias.detach() if bias is not None else None)
    return operands(cache, "linear", [weight] + ([bias] if bias is not None else []), build)


def conv_operands(cache, spec, weights):
    """per conv layer i >= 1: (wk bf16 [Cout, k*Cin], [wt_0, wt_1] for the data-gradient GEMMs)"""
    def build(prev):
        out = [None]
        for i in range(1, len(spec)):
            out.append(_be().conv_pack(weights[i].detach().contiguous(), spec[i][2], True, out=prev[i] if prev is not None else None))
        return out
    return operands(cache, "conv", list(weights[1:]), build)


def posconv_operands(cache, pos_g, pos_v):
    def build(prev):
        return _be().posconv_pack(pos_g.detach().contiguous(), pos_v.detach().contiguous(), True, out=prev)
    return operands(cache, "posconv", [pos_g, pos_v], build)


def quantizer_operands(cache, w, vars_):
    """(fp32 weight, bf16x3 split of it (B side), bf16 weight, fp32 codebook [G*V, vd], bf16 codebook)"""
    def build(prev):
        be = _be()
        w32 = w.detach().contiguous().float()
        v2 = vars_.detach().reshape(-1, vars_.shape[-1]).contiguous().float()
        p = prev if prev is not None else (None,) * 5
        return (w32, be.split3(w32, True, out=p[1]), be.cast(w32, BF16, out=p[2]), v2, be.cast(v2, BF16, out=p[4]))
    return operands(cache, "vq", [w, vars_], build)


def _bf16(t):
    """bf16 contiguous copy of a tensor through the cast kernel (parameters are fp32 masters)"""
    t = t.detach()
    if t.dtype == BF16:
        return t.contiguous()
    return _be().cast(t.contiguous().float(), BF16)


# =================================================================================================
# Linear  (nn.Linear / eight_mile Dense: wav2vec2.py:932,950,951,762)
# =================================================================================================
def _pad_rows8(t, n8):
    """[N, ...] -> [n8, ...] with zero rows appended (the tcgen05 GEMM wants N % 8 == 0: 16-byte TMA strides / stores)"""
    out = torch.zeros((n8,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    out[:t.shape[0]] = t
    return out


class LinearFn(torch.autograd.Function):
    """An output width that is not a multiple of 8 (e.g. a CTC head over a 29- or 37-symbol vocabulary, train.py's
    `len(vocab)`) runs zero-padded to the next multiple and is sliced back: the reference works for any width."""

    @staticmethod
    def forward(ctx, x, weight, bias, out_f32, cache=None):
        be = _be()
        shp = x.shape
        x2 = x.detach().reshape(-1, shp[-1])
        xb = _bf16(x2)
        N = weight.shape[0]
        wb, bvec = linear_operands(cache, weight, bias)
        N8 = wb.shape[0]
        out = _empty((x2.shape[0], N8), F32 if out_f32 else BF16, x, dynamic=True)
        be.gemm(G.linear_fwd(xb, wb, out, bvec, c_dtype=OUT_F32 if out_f32 else OUT_BF16))
        ctx.saved = (xb, wb, x.dtype, shp, bias is not None, N, ops.grad_key(weight), ops.grad_key(bias))
        if N8 != N:
            out = out[:, :N]
        return out.reshape(*shp[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        be = _be()
        xb, wb, xdtype, shp, has_bias, N, wkey, bkey = _take_saved(ctx)
        N8, K = wb.shape
        dy2 = dy.reshape(-1, N)
        if N8 != N:
            dyp = torch.zeros((dy2.shape[0], N8), dtype=dy2.dtype, device=dy2.device)
            dyp[:, :N] = dy2
            dy2 = dyp
        dyb = _bf16(dy2)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _empty(xb.shape, xdtype, xb, dynamic=True)
            be.gemm(G.linear_dgrad(dyb, wb, dx, c_dtype=OUT_F32 if xdtype == F32 else OUT_BF16))
            dx = dx.view(shp)
        if ctx.needs_input_grad[1]:
            dw = _grad_zeros(wkey, (N8, K), xb) if N8 == N else _zeros((N8, K), F32, xb)
            be.gemm(G.linear_wgrad(dyb, xb, dw))
            dw = dw[:N]
        if has_bias and ctx.needs_input_grad[2]:
            db = be.colsum(dyb, out=_grad_zeros(bkey, (N8,), xb) if N8 == N else None)[:N]
        return dx, dw, db, None, None


def linear(x, weight, bias=None, out_f32=False, cache=None):
    return LinearFn.apply(x, weight, bias, out_f32, cache)


# =================================================================================================
# LayerNorm  (nn.LayerNorm: wav2vec2.py:904,930 / :679,701)
# =================================================================================================
class LayerNormFn(torch.autograd.Function):
    """y = LN(x) in bf16 and, optionally, a second fp32 copy of the same values (for the quantizer branch)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, want_f32):
        be = _be()
        xb = _bf16(x)
        y, yf, s, mean, rstd = be.layernorm_fwd(xb, gamma.detach(), beta.detach(), eps, want_f32=want_f32)
        ctx.saved = (s, mean, rstd, gamma.detach(), x.dtype, ops.grad_key(gamma), ops.grad_key(beta))
        if want_f32:
            return y, yf
        return y

    @staticmethod
    def backward(ctx, dy, dyf=None):
        be = _be()
        s, mean, rstd, gamma, xdtype, gkey, bkey = _take_saved(ctx)
        if dy is None:
            dy = _zeros(s.shape, BF16, s)
        C = s.shape[-1]
        ds, _, dg, db, _ = be.layernorm_bwd(_bf16(dy), s, mean, rstd, gamma,
                                            dy_f32=dyf.contiguous() if dyf is not None else None,
                                            dg_out=_grad_zeros(gkey, (C,), s), db_out=_grad_zeros(bkey, (C,), s))
        return ds.to(xdtype), dg, db, None, None


def layer_norm(x, gamma, beta, eps, want_f32=False):
    return LayerNormFn.apply(x, gamma, beta, eps, want_f32)


# =================================================================================================
# Dropout (nn.Dropout on activations: wav2vec2.py:934-935, 713)
# =================================================================================================
class DropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p, seed):
        src = step_seed(x)
        ctx.cfg = (p, seed, src)
        _be().set_seed_source(src)
        out = _be().dropout(x.detach().contiguous(), p, seed)
        _be().set_seed_source(None)
        return out

    @staticmethod
    def backward(ctx, dy):
        p, seed, src = ctx.cfg
        _be().set_seed_source(src)
        dx = _be().dropout(dy.contiguous(), p, seed)
        _be().set_seed_source(None)
        return dx, None, None


def dropout(x, p, training):
    if not training or p <= 0.0:
        return x
    return DropoutFn.apply(x, float(p), next_seed())


# =================================================================================================
# time-mask plumbing (wav2vec2.py:939, 946, 381, 717, 721)
# =================================================================================================
def span_mask_draw(like, B, T, p_start, mask_length, R_max):
    """device-side `create_mask` (reference wav2vec2.py:189-216; csrc/draws.cu): -> (padded masked-row list int32
    [R_max + 1] with the count in its last element, mask uint8 [B,T]).  The per-step seed is a device word from torch's
    CUDA generator (`step_seed`), so the draw follows torch.manual_seed() and changes on every CUDA-graph replay."""
    with torch.no_grad():
        return _be().span_mask_draw(next_seed(), step_seed(like), B, T, p_start, mask_length, R_max, like.device)


def negatives_draw(rows, B, K):
    """device-side `Sampler.negatives` indices (reference wav2vec2.py:955-976) for the padded row list of
    `span_mask_draw` -> int32 [R_max * K]"""
    with torch.no_grad():
        return _be().negatives_draw(next_seed(), step_seed(rows), rows, B, K)


class RowsSetFn(torch.autograd.Function):
    """features[time_mask] = mask_emb   (rows of a [B,T,C] bf16 tensor given flat row indices)"""

    @staticmethod
    def forward(ctx, x, idx, vec):
        out = x.detach().clone()
        _be().rows_set(out.view(-1, out.shape[-1]), idx, vec.detach().float().contiguous())
        ctx.idx = idx
        ctx.vkey = ops.grad_key(vec)
        return out

    @staticmethod
    def backward(ctx, dy):
        dx = dy.clone().contiguous()
        C = dx.shape[-1]
        dvec = _be().rows_set_bwd(dx.view(-1, C), ctx.idx, out=_grad_zeros(ctx.vkey, (C,), dx))
        return dx, None, dvec


class RowsGatherFn(torch.autograd.Function):
    """x[time_mask] as rows: [n, C] gathered from a [B*T, C] view (fp32 or bf16 in, fp32 out)"""

    @staticmethod
    def forward(ctx, x, idx):
        x2 = x.detach().reshape(-1, x.shape[-1]).contiguous()
        ctx.cfg = (idx, x.shape, x.dtype)
        return _be().rows_gather(x2, idx, F32)

    @staticmethod
    def backward(ctx, dy):
        idx, shp, dtype = ctx.cfg
        n_rows = 1
        for d in shp[:-1]:
            n_rows *= d
        dx = _be().rows_scatter(dy.contiguous().float(), idx, n_rows, dtype)
        return dx.view(shp), None


class MaskApplyFn(torch.autograd.Function):
    """zero padded frames (`x[~pad_mask] = 0`, wav2vec2.py:632) and masked channels (:721); self-adjoint"""

    @staticmethod
    def forward(ctx, x, row_keep, chan_zero):
        out = x.detach().clone()
        _be().mask_apply(out, row_keep, chan_zero)
        ctx.cfg = (row_keep, chan_zero)
        return out

    @staticmethod
    def backward(ctx, dy):
        dx = dy.clone().contiguous()
        _be().mask_apply(dx, *ctx.cfg)
        return dx, None, None


# =================================================================================================
# conv feature encoder (wav2vec2.py:399-456)
# =================================================================================================
class ConvFeatureFn(torch.autograd.Function):
    """x fp32 [B,L] -> bf16 [B,T,C] channels-last.  Layer 0 = fused conv+GroupNorm+GELU kernels (csrc/conv0.cu),
    layers 1.. = implicit-GEMM tcgen05 convs with GELU epilogues (zero-copy im2col through overlapping TMA rows)."""

    @staticmethod
    def forward(ctx, x, spec, cache, gn_w, gn_b, *weights):
        be = _be()
        x = x.detach().contiguous().float()
        (c0, k0, s0) = spec[0]
        w0 = weights[0].detach().reshape(c0, k0).contiguous()
        gw, gb = gn_w.detach(), gn_b.detach()
        need_grad = any(ctx.needs_input_grad)
        mean, rstd, mom = be.conv0_stats(x, w0, k0, s0, 1e-5)
        a = be.conv0_fwd(x, w0, gw, gb, mean, rstd, k0, s0)
        acts, zs, wts = [a], [None], [None]
        packed = conv_operands(cache, spec, weights)
        for i in range(1, len(spec)):
            (c, k, s) = spec[i]
            B, Lin, Cin = a.shape
            Lout = (Lin - k) // s + 1
            wk, wt = packed[i]
            y = _empty((B, Lout, c), BF16, a)
            z = _empty((B, Lout, c), F16, a) if need_grad else None
            be.gemm(G.conv_fwd(a, wk, y, k, s, z_out=z))
            acts.append(y)
            zs.append(z)
            wts.append(wt)
            a = y
        if need_grad:
            # never keep a RETURNED tensor object in ctx: its grad_fn is this node, and the reference cycle would hold
            # every saved activation until Python's cyclic GC runs (a detached alias shares the storage, not the cycle)
            acts[-1] = a.detach()
            keys = [ops.grad_key(w) for w in weights] + [ops.grad_key(gn_w), ops.grad_key(gn_b)]
            ctx.saved = (x, w0, gw, gb, mean, rstd, mom, acts, zs, wts, spec, keys)
        return a

    @staticmethod
    def backward(ctx, dy):
        be = _be()
        x, w0, gw, gb, mean, rstd, mom, acts, zs, wts, spec, keys = _take_saved(ctx)
        n = len(spec)
        grads = [None] * n
        if n > 1:
            dz = be.mul_dgelu(_bf16(dy), zs[n - 1])  # zs hold gelu'(pre-activation) in fp16 (ACT_GELU_DZ)
            for i in range(n - 1, 0, -1):
                (c, k, s) = spec[i]
                a_prev = acts[i - 1]
                Cin = a_prev.shape[2]
                dwk = _zeros((c, k * Cin), F32, a_prev)
                be.gemm(G.conv_wgrad(dz, a_prev, dwk, k, s))
                grads[i] = be.conv_unpack(dwk, Cin, k, out=_grad_empty(keys[i], (c, Cin, k), a_prev))
                dprev = _empty(a_prev.shape, BF16, a_prev)
                for p in range(s):
                    be.gemm(G.conv_dgrad(dz, wts[i][p], dprev, k, s, p, aux=zs[i - 1] if i > 1 else None))
                dz = dprev  # for i == 1 this is dL/d(a0), the gradient w.r.t. layer 0's GELU output
                acts[i] = zs[i] = None
            da0 = dz
        else:
            da0 = _bf16(dy)
        (c0, k0, s0) = spec[0]
        dest = (_grad_empty(keys[0], (c0, k0), x), _grad_empty(keys[n], (c0,), x), _grad_empty(keys[n + 1], (c0,), x))
        dw0, dg, db = be.conv0_bwd(x, w0, gw, gb, mean, rstd, mom, k0, s0, da0, out=dest)
        grads[0] = dw0.view(c0, 1, k0)
        return (None, None, None, dg, db, *grads)


# =================================================================================================
# transformer encoder with positional conv (wav2vec2.py:579-646 + eight_mile TransformerEncoderStack)
# =================================================================================================
def _prepare_layer_weights(be, arena, lw, per_layer):
    """one launch: every transformer GEMM weight fp32 -> bf16 into a persistent arena (w_Q|w_K|w_V land in one fused
    [3D,D] operand, their biases in one fp32 [3D] vector).  The buffers and the device-side pointer table persist; the
    values are rewritten only when a parameter changed (see `operands`)."""
    nl = len(lw) // per_layer
    D = lw[0].shape[1]
    F_ = lw[10].shape[0]
    dev = lw[0].device
    sig = (nl, D, F_, dev)
    if arena.get("shape") != sig:
        per = 4 * D * D + 2 * F_ * D
        wbuf = torch.empty(nl * per, dtype=BF16, device=dev)
        bbuf = torch.empty(nl * 3 * D, dtype=F32, device=dev)
        views, dsts = [], []
        for li in range(nl):
            o = li * per
            wqkv = wbuf[o:o + 3 * D * D].view(3 * D, D)
            wo = wbuf[o + 3 * D * D:o + 4 * D * D].view(D, D)
            w1 = wbuf[o + 4 * D * D:o + 4 * D * D + F_ * D].view(F_, D)
            w2 = wbuf[o + 4 * D * D + F_ * D:o + per].view(D, F_)
            bqkv = bbuf[li * 3 * D:(li + 1) * 3 * D]
            views.append((wqkv, wo, w1, w2, bqkv))
            dsts.append([wqkv[0:D], wqkv[D:2 * D], wqkv[2 * D:], wo, w1, w2, bqkv[0:D], bqkv[D:2 * D], bqkv[2 * D:]])
        arena.update(shape=sig, wbuf=wbuf, bbuf=bbuf, views=views, dsts=dsts, cache={})
    pairs = []
    for li in range(nl):
        (wq, bq, wk, bk, wv, bv, wo, _bo, _g2, _b2, w1, _b1, w2, _bb2, _g1, _b1l) = lw[li * per_layer:(li + 1) * per_layer]
        srcs = [wq, wk, wv, wo, w1, w2, bq, bk, bv]
        pairs += [(sv.detach(), d) for sv, d in zip(srcs, arena["dsts"][li])]
    ver = tuple((sv.data_ptr(), sv._version) for sv, _ in pairs)
    if arena.get("ver") != ver:
        be.cast_multi(pairs, arena["cache"])
        arena["ver"] = ver
    return arena["views"]


def encoder_operands(arena, pos_g, pos_v, lw, per_layer, front=True):
    """refresh (if stale) the transformer stack's bf16 weight arena and the packed positional conv: what
    AudioTransformerEncoder calls before it replays its CUDA-graph segments"""
    if front and pos_v is not None:
        posconv_operands(arena, pos_g, pos_v)
    if len(lw):
        _prepare_layer_weights(_be(), arena, lw, per_layer)


class EncoderFn(torch.autograd.Function):
    """AudioTransformerEncoder.extract_features:  x (+pad zeroing) -> x + gelu(pos_conv(x)) -> LN -> dropout ->
    num_layers x post-LN transformer layer.  Per layer the parameters arrive as
    (w_Q, b_Q, w_K, b_K, w_V, b_V, w_O, b_O, ln2_g, ln2_b, w_1 [F,D], b_1, w_2 [D,F], b_2, ln1_g, ln1_b)."""

    PER_LAYER = 16

    @staticmethod
    def forward(ctx, x, cfg, row_keep, pos_g, pos_v, pos_b, ln_g, ln_b, *lw):
        be = _be()
        H, groups, pdrop, training, active = cfg["num_heads"], cfg["groups"], cfg["pdrop"], cfg["training"], cfg["active"]
        PL = EncoderFn.PER_LAYER
        p = pdrop if training else 0.0
        x = _bf16(x)
        B, T, D = x.shape
        M = B * T
        if D % (8 * groups) != 0 or D != H * 64:
            raise ValueError(f"d_model={D}, heads={H}: this build needs d_model % {8 * groups} == 0 and d_k == 64 "
                             "(wav2vec2 base and large both use 64)")
        k = pos_v.shape[-1] if pos_v is not None else 0
        pad_l = k // 2 - 1 if k % 2 == 0 else k // 2
        if row_keep is not None and cfg.get("front", True):  # `x[~pad_mask] = 0` (wav2vec2.py:632): encoder input only
            x = x.clone()
            be.mask_apply(x, row_keep, None)
        need_grad = any(ctx.needs_input_grad)
        seed_src = step_seed(x) if p > 0 else None
        be.set_seed_source(seed_src)
        front = cfg.get("front", True)  # False: a later slice of the stack (no positional conv / LayerNorm in front)
        if front:
            pg, pv = pos_g.detach().contiguous(), pos_v.detach().contiguous()
            wp, wpt, norm2 = posconv_operands(cfg["arena"], pos_g, pos_v)
            s0 = _empty(x.shape, BF16, x)
            z0 = _empty(x.shape, F16, x)
            be.gemm(G.posconv_fwd(x, wp, s0, pos_b.detach(), groups, k, pad_l, z_out=z0))
            seed0 = next_seed() if p > 0 else 0
            h, _, _, mean0, rstd0 = be.layernorm_fwd(s0, ln_g.detach(), ln_b.detach(), 1e-5, p_y=p, seed_y=seed0)
        else:
            pg = pv = wpt = norm2 = s0 = z0 = mean0 = rstd0 = None
            seed0 = 0
            h = x
        Tp = (T + 7) // 8 * 8
        scale = 1.0 / math.sqrt(D // H)
        wviews = _prepare_layer_weights(be, cfg["arena"], lw, PL) if len(lw) else []
        layers = []
        for li in range(len(lw) // PL):
            if not active[li]:
                layers.append(None)
                continue
            (_wq, _bq, _wk, _bk, _wv, _bv, _wo, bo, g2, b2, _w1, b1, _w2, bb2, g1, b1ln) = (
                t.detach() for t in lw[li * PL:(li + 1) * PL])
            wqkv_b, wo_b, w1_b, w2_b, bqkv = wviews[li]
            F_ = w1_b.shape[0]
            xin = h
            x2d = xin.view(M, D)
            qkv = _empty((B, T, 3 * D), BF16, x)
            be.gemm(G.linear_fwd(x2d, wqkv_b, qkv.view(M, 3 * D), bqkv))
            seed_a = next_seed() if p > 0 else 0
            ctxv, lse = be.attn_fwd(qkv, H, scale, row_keep, p, seed_a)
            a = _empty((M, D), BF16, x)
            be.gemm(G.linear_fwd(ctxv.view(M, D), wo_b, a, bo))
            seed1 = next_seed() if p > 0 else 0
            x1, _, s1, mean2, rstd2 = be.layernorm_fwd(x2d, g2, b2, 1e-6, h=a, p_h=p, seed_h=seed1)
            z1 = _empty((M, F_), F16, x) if need_grad else None
            hid = _empty((M, F_), BF16, x)
            be.gemm(G.linear_fwd(x1, w1_b, hid, b1, act=ACT_GELU_DZ if need_grad else ACT_GELU, z_out=z1 if need_grad else None))
            f = _empty((M, D), BF16, x)
            be.gemm(G.linear_fwd(hid, w2_b, f, bb2))
            seed2 = next_seed() if p > 0 else 0
            x2, _, s2, mean1, rstd1 = be.layernorm_fwd(x1, g1, b1ln, 1e-6, h=f, p_h=p, seed_h=seed2)
            h = x2.view(B, T, D)
            if need_grad:
                layers.append(dict(xin=x2d, qkv=qkv, lse=lse, ctx=ctxv, s1=s1, mean2=mean2, rstd2=rstd2, x1=x1,
                                   z1=z1, hid=hid, s2=s2, mean1=mean1, rstd1=rstd1, seeds=(seed_a, seed1, seed2),
                                   w=(wqkv_b, wo_b, w1_b, w2_b), ln=(g2, g1), key=ops.grad_key(lw[li * PL])))
        if need_grad:
            ctx.saved = dict(x=x, s0=s0, z0=z0, mean0=mean0, rstd0=rstd0, seed0=seed0, front=front,
                             ln_g=ln_g.detach() if front else None,
                             fkeys=[ops.grad_key(t) for t in (pos_g, pos_v, pos_b, ln_g, ln_b)] if front else None,
                             lkeys=[ops.grad_key(lw[li * PL]) for li in range(len(lw) // PL)], F_=lw[10].shape[0] if len(lw) else 0,
                             pg=pg, pv=pv, norm2=norm2, wpt=wpt, layers=layers, row_keep=row_keep, p=p, H=H,
                             groups=groups, k=k, pad_l=pad_l, Tp=Tp, scale=scale, nlw=len(lw), seed_src=seed_src)
        be.set_seed_source(None)
        return h

    @staticmethod
    def _layer_views(zbuf, D, F_):
        """one layer's gradient block [dWqkv | dWo | dW1 | dW2 | (dg1, db1ln, db2) | (dg2, db2ln, dbo) | db1 | dbqkv] as
        (the 16 per-parameter gradients in `PER_LAYER` order, the accumulators the kernels write)"""
        sizes = [3 * D * D, D * D, F_ * D, D * F_, 3 * D, 3 * D, F_, 3 * D]
        parts, o = [], 0
        for n_ in sizes:
            parts.append(zbuf[o:o + n_])
            o += n_
        dwqkv, dwo, dw1, dw2 = parts[0].view(3 * D, D), parts[1].view(D, D), parts[2].view(F_, D), parts[3].view(D, F_)
        acc1, acc2, db1, dbqkv = parts[4].view(3, D), parts[5].view(3, D), parts[6], parts[7]
        grads = [dwqkv[0:D], dbqkv[0:D], dwqkv[D:2 * D], dbqkv[D:2 * D], dwqkv[2 * D:], dbqkv[2 * D:], dwo, acc2[2],
                 acc2[0], acc2[1], dw1, db1, dw2, acc1[2], acc1[0], acc1[1]]
        return grads, (dwqkv, dwo, dw1, dw2, acc1, acc2, db1, dbqkv)

    @staticmethod
    def backward(ctx, dh):
        be = _be()
        sv = _take_saved(ctx)
        PL = EncoderFn.PER_LAYER
        p, H, Tp, scale = sv["p"], sv["H"], sv["Tp"], sv["scale"]
        x = sv["x"]
        B, T, D = x.shape
        M = B * T
        dcur = _bf16(dh).view(M, D)
        be.set_seed_source(sv["seed_src"])
        lgrads = [None] * sv["nlw"]
        # The 4 weight-gradient GEMMs of every layer (dW = dY^T X) are deferred and issued as ONE grouped persistent launch
        # when the stack's data-gradient chain is done (ops.gemm_group): per-launch set-up, first-load latency and the
        # exposed last epilogue are paid once instead of 48 times, the group needs no split-K (plain stores: no zero fill
        # of dW), and the contraction (B*T tokens) is long enough to run at the conv-wgrad rate.  Costs ~62 MB per layer of
        # gradients kept alive until then.
        deferred = []
        for li in range(len(sv["layers"]) - 1, -1, -1):
            L = sv["layers"][li]
            if L is None:
                # LayerDrop skipped this layer.  Under the data-parallel wrapper its gradient block still takes part in
                # the all-reduce (other ranks may have run the layer): hand out the zeroed block as this rank's gradients
                F0 = sv["F_"]
                zb = ops.grad_arena_take(sv["lkeys"][li], 4 * D * D + 2 * F0 * D + 9 * D + F0)
                if zb is not None:
                    lgrads[li * PL:(li + 1) * PL] = EncoderFn._layer_views(zb, D, F0)[0]
                continue
            wqkv_b, wo_b, w1_b, w2_b = L["w"]
            g2, g1 = L["ln"]
            seed_a, seed1, seed2 = L["seeds"]
            F_ = w1_b.shape[0]
            # every fp32 accumulator of this layer's backward comes from ONE zero-filled buffer (one fill launch)
            sizes = [3 * D * D, D * D, F_ * D, D * F_, 3 * D, 3 * D, F_, 3 * D]
            # data-parallel runs: the block lives in the wrapper's gradient arena (parallel.py) and is reduced in place
            zbuf = ops.grad_arena_take(L["key"], sum(sizes), zero=False)
            if zbuf is None:
                zbuf = _empty((sum(sizes),), F32, x)
            zbuf[sum(sizes[:4]):].zero_()  # only the small bias / LayerNorm accumulators are accumulated into
            _, (dwqkv, dwo, dw1, dw2, acc1, acc2, db1, dbqkv) = EncoderFn._layer_views(zbuf, D, F_)
            # ---- ln1( x1 + drop(ffn) )
            ds2, df, dg1, db1ln, dbias2 = be.layernorm_bwd(dcur, L["s2"], L["mean1"], L["rstd1"], g1, want_dh=p > 0,
                                                           p_h=p, seed_h=seed2, want_dbias=True, acc=acc1)
            if df is None:
                df = ds2
            deferred.append(G.linear_wgrad_grouped(df, L["hid"], dw2))
            dz1 = _empty((M, F_), BF16, x)
            # (the bias gradient db1 = column sums of dz1 comes out of the same epilogue: no separate pass over dz1)
            be.gemm(G.linear_dgrad(df, w2_b, dz1, aux=L["z1"], aux_mode=AUX_MUL, colsum=db1))
            deferred.append(G.linear_wgrad_grouped(dz1, L["x1"], dw1))
            dx1 = _empty((M, D), BF16, x)
            be.gemm(G.linear_dgrad(dz1, w1_b, dx1, aux=ds2, aux_mode=AUX_ADD))
            # ---- ln2( x + drop(attn) )
            ds1, da, dg2, db2ln, dbo = be.layernorm_bwd(dx1, L["s1"], L["mean2"], L["rstd2"], g2, want_dh=p > 0,
                                                        p_h=p, seed_h=seed1, want_dbias=True, acc=acc2)
            if da is None:
                da = ds1
            deferred.append(G.linear_wgrad_grouped(da, L["ctx"].view(M, D), dwo))
            dctx = _empty((B, T, D), BF16, x)
            be.gemm(G.linear_dgrad(da, wo_b, dctx.view(M, D)))
            # (the fused QKV bias gradient dbqkv = column sums of dqkv leaves the attention backward kernels with it)
            dqkv = be.attn_bwd(L["qkv"], L["ctx"], dctx, L["lse"], H, scale, sv["row_keep"], p, seed_a, dbias=dbqkv)
            dqkv2 = dqkv.view(M, 3 * D)
            deferred.append(G.linear_wgrad_grouped(dqkv2, L["xin"], dwqkv))
            dxin = _empty((M, D), BF16, x)
            be.gemm(G.linear_dgrad(dqkv2, wqkv_b, dxin, aux=ds1, aux_mode=AUX_ADD))
            dcur = dxin
            lgrads[li * PL:(li + 1) * PL] = [dwqkv[0:D], dbqkv[0:D], dwqkv[D:2 * D], dbqkv[D:2 * D], dwqkv[2 * D:],
                                             dbqkv[2 * D:], dwo, dbo, dg2, db2ln, dw1, db1, dw2, dbias2, dg1, db1ln]
            # (the layer's activations that the deferred weight gradients read stay referenced by `deferred`)
            sv["layers"][li] = None
        if deferred:
            be.gemm_group(deferred)
            deferred = None
        if not sv["front"]:
            be.set_seed_source(None)
            return (dcur.view(B, T, D), None, None, None, None, None, None, None, *lgrads)
        # ---- front: LN(+dropout) <- x + gelu(pos_conv(x))
        kg, kv, kb_, klg, klb = sv["fkeys"]
        ds0, _, dlg, dlb, _ = be.layernorm_bwd(dcur.view(B, T, D), sv["s0"], sv["mean0"], sv["rstd0"], sv["ln_g"],
                                               p_y=p, seed_y=sv["seed0"], dg_out=_grad_zeros(klg, (D,), x),
                                               db_out=_grad_zeros(klb, (D,), x))
        dz0 = be.mul_dgelu(ds0, sv["z0"])
        dpos_b = be.colsum(dz0, out=_grad_zeros(kb_, (D,), x))
        groups, k, pad_l = sv["groups"], sv["k"], sv["pad_l"]
        dwp = _empty((groups, k * 64, 64), F32, x)
        be.gemm(G.posconv_wgrad(dz0, x, dwp, groups, k, pad_l))
        dpos_v, dpos_g = be.posconv_wn_bwd(dwp, sv["pg"], sv["pv"], sv["norm2"],
                                           out=(_grad_empty(kv, tuple(sv["pv"].shape), x), _grad_empty(kg, tuple(sv["pg"].shape), x)))
        dx = _empty((B, T, D), BF16, x)
        be.gemm(G.posconv_dgrad(dz0, sv["wpt"], dx, groups, k, pad_l, aux=ds0))
        if sv["row_keep"] is not None:
            be.mask_apply(dx, sv["row_keep"], None)
        be.set_seed_source(None)
        return (dx, None, None, dpos_g, dpos_v, dpos_b, dlg, dlb, *lgrads)


# =================================================================================================
# Gumbel vector quantizer (wav2vec2.py:547-576)
# =================================================================================================
class QuantizerFn(torch.autograd.Function):
    """n_valid: optional int32 device scalar — only the first n_valid of the B*Tm rows exist, the rest is the padding of
    a worst-case-length (static-shape, CUDA-graph-replayable) row list."""

    keep_logits = False  # parity tests: expose the fp32 logits of the last call (GumbelVectorQuantizer.keep_logits)
    last_logits = None

    @staticmethod
    def forward(ctx, y, w, b, vars_, G_, tau, noise, n_valid=None, cache=None):
        be = _be()
        Bq, Tm, Cin = y.shape
        R = Bq * Tm
        y2 = y.detach().reshape(R, Cin).contiguous().float()
        w32, w_split, w_bf, v2, v_bf = quantizer_operands(cache, w, vars_)
        # logits with fp32-accurate products on the tensor cores: bf16x3 split along K (csrc/misc.cu)
        z = _empty((R, w.shape[0]), F32, y, dynamic=True)
        be.gemm(G.linear_fwd(be.split3(y2, False), w_split, z, b.detach(), c_dtype=OUT_F32))
        q, qb, kidx, avg, ppl = be.vq_fwd(z, noise, float(tau), v2, G_, n_valid=n_valid)
        # outputs are kept as detached aliases (no ctx <-> output reference cycle)
        ctx.saved = (y2, w32, v2, z, noise, kidx.detach(), avg, ppl.detach(), G_, float(tau), y.shape, vars_.shape,
                     (ops.grad_key(w), ops.grad_key(b), ops.grad_key(vars_)), n_valid, w_bf, v_bf)
        ctx.mark_non_differentiable(kidx)
        QuantizerFn.last_logits = z if QuantizerFn.keep_logits else None
        return q.view(Bq, Tm, -1), ppl, kidx

    @staticmethod
    def backward(ctx, dq, dppl, _dk):
        be = _be()
        y2, w32, v2, z, noise, kidx, avg, ppl, G_, tau, yshape, vshape, (wkey, bkey, vkey), n_valid, w_bf, v_bf = _take_saved(ctx)
        R = y2.shape[0]
        vd = v2.shape[1]
        if dq is None:
            dq = _zeros((R, G_ * vd), F32, y2)
        dq2 = dq.reshape(R, G_ * vd).contiguous().float()
        if dppl is None:
            dppl = _zeros((), F32, y2)
        a_dot = None
        if noise is not None:
            a_dot = _empty((R, v2.shape[0]), F32, y2, dynamic=True)
            be.gemm(G.vq_codebook_dots(_bf16(dq2), v_bf, a_dot, G_))
        dz, dvars = be.vq_bwd(z, noise, tau, G_, vd, a_dot, dq2, kidx, avg, ppl, dppl.contiguous().float(),
                              dvars_out=_grad_zeros(vkey, tuple(v2.shape), y2), n_valid=n_valid)
        db = be.colsum(dz, out=_grad_zeros(bkey, (dz.shape[-1],), y2))
        dw = _grad_zeros(wkey, tuple(w32.shape), y2)
        be.gemm(G.linear_wgrad(dz, _bf16(y2), dw))
        dy = _empty(y2.shape, F32, y2, dynamic=True)
        be.gemm(G.linear_dgrad(dz, w_bf, dy, c_dtype=OUT_F32))
        return dy.view(yshape), dw, db, dvars.view(vshape), None, None, None, None, None


# =================================================================================================
# contrastive loss (wav2vec2.py:377-392)
# =================================================================================================
class ContrastiveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, idx, ppl, n_vars, xe_w, div_w, n_valid=None):
        be = _be()
        x2 = x.detach().reshape(-1, x.shape[-1]).contiguous().float()
        y2 = y.detach().reshape(-1, y.shape[-1]).contiguous().float()
        loss, ce, saved = be.contrastive_fwd(x2, y2, idx, ppl.detach().float() if ppl is not None else None,
                                             float(n_vars), float(xe_w), float(div_w), n_valid=n_valid)
        ctx.saved = (x2, y2, idx, saved, x.shape, y.shape, float(n_vars), float(xe_w), float(div_w), ppl is not None,
                     n_valid)
        return loss, ce

    @staticmethod
    def backward(ctx, dloss, dce_extra):
        be = _be()
        x2, y2, idx, saved, xs, ys, n_vars, xe_w, div_w, has_ppl, n_valid = _take_saved(ctx)
        dce = dloss * xe_w
        if dce_extra is not None:
            dce = dce + dce_extra
        dx, dy = be.contrastive_bwd(x2, y2, idx, saved, dce.contiguous().float(), n_valid=n_valid)
        dppl = (-div_w / n_vars) * dloss if has_ppl else None
        return dx.view(xs), dy.view(ys), None, dppl, None, None, None, None


# =================================================================================================
# log-softmax head (wav2vec2.py:770)
# =================================================================================================
class LogSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = _be().log_softmax_fwd(x.detach().contiguous().float())
        ctx.saved = y.detach()  # alias, not the returned object (no reference cycle through grad_fn)
        return y

    @staticmethod
    def backward(ctx, dy):
        y = _take_saved(ctx)
        dx = _be().log_softmax_bwd(dy.float(), y)  # strided dy (e.g. CTC's [T,B,V] transposed back) is consumed as is
        return dx.float()
