"""TEST INFRASTRUCTURE — pure-PyTorch emulation of the C ABI in include/audio8_b200.h.

CPU tests inject `EmuBackend` with `audio8_b200.ops.set_backend(...)` to exercise the host-side orchestration
(GEMM descriptors, autograd wiring, module logic) without a GPU; GPU tests use the same functions as the
per-kernel reference.  It mirrors the ABI semantics literally (TMA boxes with zero fill, tile loops, epilogue
order), not efficiently.  The product package never imports this file.
"""
import math

import torch

from audio8_b200 import _lib

BM, BK = 128, 64


def _flat(t):
    assert t.is_contiguous()
    return t.reshape(-1)


def _box(op, coords, box0, box1):
    """TMA tiled load of a {box0, box1, 1, 1} box at `coords`, out-of-range elements read as zero -> [box1, box0] fp32"""
    flat = _flat(op.t)
    d, s = op.dims, op.strides
    i0 = coords[0] + torch.arange(box0)
    i1 = coords[1] + torch.arange(box1)
    ok = ((i0 >= 0) & (i0 < d[0]))[None, :] & ((i1 >= 0) & (i1 < d[1]))[:, None]
    if not (0 <= coords[2] < d[2] and 0 <= coords[3] < d[3]):
        return torch.zeros(box1, box0)
    idx = op.offset + i0[None, :] + i1[:, None] * s[0] + coords[2] * s[1] + coords[3] * s[2]
    idx = idx.clamp(0, flat.numel() - 1)
    return torch.where(ok, flat[idx].float(), torch.zeros(()))


def _coords(op, kin, kbatch, r, lo, hi):
    return [op.base[d] + op.ck[d] * kin + op.cb[d] * kbatch + op.cr[d] * r + op.cl[d] * lo + op.ch[d] * hi
            for d in range(4)]


def _tile(op, rows, row0, kin, kbatch, lo, hi):
    """operand tile [rows, 64] (row-major over the tile's M/N rows, 64 k-elements)"""
    if op.major == _lib.MAJOR_K:
        return _box(op, _coords(op, kin, kbatch, row0, lo, hi), BK, rows)
    parts = []
    for at in range(rows // 64):
        blk = _box(op, _coords(op, kin, kbatch, row0 // 64 + at, lo, hi), 64, BK)  # [64 k-rows, 64 mn]
        parts.append(blk.t())
    return torch.cat(parts, 0)


def gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * 0.7071067811865476))


def gelu_grad(x):
    return 0.5 * (1.0 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327


def emu_gemm(g):
    bn = g.block_n or (64 if g.N <= 64 else (128 if g.N <= 128 else 256))
    c = _flat(g.c)
    z = _flat(g.z_out) if g.z_out is not None else None
    aux = _flat(g.aux) if g.aux is not None else None
    for hi in range(g.hi_count):
        for lo in range(g.lo_count):
            for mt in range(math.ceil(g.M / BM)):
                for nt in range(math.ceil(g.N / bn)):
                    m0, n0 = mt * BM, nt * bn
                    acc = torch.zeros(BM, bn)
                    for kb in range(g.k_blocks):
                        kin, kbatch = kb % g.k_inner, kb // g.k_inner
                        a = _tile(g.a, BM, m0, kin, kbatch, lo, hi)
                        b = _tile(g.b, bn, n0, kin, kbatch, lo, hi)
                        acc += a @ b.t()
                    mv, nv = min(BM, g.M - m0), min(bn, g.N - n0)
                    v = acc[:mv, :nv] * g.alpha
                    if g.bias is not None:
                        v = v + g.bias[lo * g.bias_stride_lo + n0: lo * g.bias_stride_lo + n0 + nv].float()[None, :]
                    rows = torch.arange(m0, m0 + mv)
                    cols = torch.arange(n0, n0 + nv)
                    off = g.c_offset + hi * g.c_stride_hi + lo * g.c_stride_lo + rows[:, None] * g.ldc + cols[None, :]
                    if z is not None:
                        z[off] = (gelu_grad(v) if g.act == _lib.ACT_GELU_DZ else v).to(z.dtype)
                    if g.act in (_lib.ACT_GELU, _lib.ACT_GELU_DZ):
                        v = gelu(v)
                    if g.aux_mode == _lib.AUX_ADD:
                        v = v + aux[off].float()
                    elif g.aux_mode == _lib.AUX_MUL_GELU_GRAD:
                        v = v * gelu_grad(aux[off].float())
                    elif g.aux_mode == _lib.AUX_MUL:
                        v = v * aux[off].float()
                    if g.c_dtype == _lib.OUT_F32_ATOMIC:
                        c[off] += v
                    else:
                        c[off] = v.to(c.dtype)
                        if getattr(g, "colsum", None) is not None:  # column sums of the output as stored
                            g.colsum[n0:n0 + nv] += c[off].float().sum(0)


class EmuBackend:
    name = "emu"

    def gemm(self, g):
        emu_gemm(g.spec() if hasattr(g, "spec") else g)

    def gemm_group(self, specs):
        for g in specs:
            self.gemm(g)


_SEED_SRC = [None]


def _drop_mask(shape, p, seed):
    if _SEED_SRC[0] is not None:
        seed = int(seed) + int(_SEED_SRC[0].reshape(-1)[0].item())
    """consistent between forward and backward within the emulation (not bit-identical to the CUDA Philox stream)"""
    if p <= 0:
        return torch.ones(shape)
    g = torch.Generator().manual_seed(int(seed) & 0x7FFFFFFF)
    return (torch.rand(shape, generator=g) >= p).float() / (1.0 - p)


def _bf(x):
    return x.to(torch.bfloat16)


def _into(out, val):
    """the ABI's `out=` destinations (zeroed accumulators / overwritten buffers) as a plain copy"""
    if out is None:
        return val
    out.copy_(val.reshape(out.shape))
    return out


class EmuOps(EmuBackend):
    """row / index / loss kernels of the ABI, emulated with fp32 PyTorch math and bf16 storage"""

    # ---- layernorm
    def set_seed_source(self, t):
        _SEED_SRC[0] = t

    def layernorm_fwd(self, x, gamma, beta, eps, h=None, p_h=0.0, seed_h=0, want_f32=False, p_y=0.0, seed_y=0):
        s = x
        if h is not None:
            s = _bf(x.float() + h.float() * _drop_mask(h.shape, p_h, seed_h))
        sf = s.float()
        mean = sf.mean(-1)
        var = sf.var(-1, unbiased=False)
        rstd = (var + eps).rsqrt()
        y = ((sf - mean[..., None]) * rstd[..., None] * gamma + beta) * _drop_mask(s.shape, p_y, seed_y)
        return _bf(y), (y.clone() if want_f32 else None), s, mean.reshape(-1), rstd.reshape(-1)

    def layernorm_bwd(self, dy, s, mean, rstd, gamma, dy_f32=None, p_y=0.0, seed_y=0, want_dh=False, p_h=0.0,
                      seed_h=0, want_dbias=False, acc=None, dg_out=None, db_out=None):
        C = s.shape[-1]
        g = dy.float() + (dy_f32 if dy_f32 is not None else 0.0)
        g = (g * _drop_mask(s.shape, p_y, seed_y)).reshape(-1, C)
        xh = (s.float().reshape(-1, C) - mean[:, None]) * rstd[:, None]
        dgamma, dbeta = (g * xh).sum(0), g.sum(0)
        gg = g * gamma
        ds = rstd[:, None] * (gg - gg.mean(-1, keepdim=True) - xh * (gg * xh).mean(-1, keepdim=True))
        ds = ds.reshape(s.shape)
        dh = None
        if want_dh:
            dh = ds * _drop_mask(s.shape, p_h, seed_h)
        dbias = None
        if want_dbias:
            dbias = (dh if dh is not None else ds).reshape(-1, C).sum(0)
        if acc is not None:
            dgamma, dbeta = _into(acc[0], dgamma), _into(acc[1], dbeta)
            dbias = _into(acc[2], dbias) if dbias is not None else None
        elif dg_out is not None and not want_dbias:
            dgamma, dbeta = _into(dg_out, dgamma), _into(db_out, dbeta)
        return _bf(ds), (_bf(dh) if dh is not None else None), dgamma, dbeta, dbias

    # ---- softmax
    def softmax_fwd(self, s, T, key_keep=None, pdrop=0.0, seed=0):
        B, H, _, Tp = s.shape
        x = s[..., :T].clone()
        if key_keep is not None:
            x = x.masked_fill(key_keep[:, None, None, :] == 0, -1e9)
        p = torch.zeros(s.shape)
        p[..., :T] = torch.softmax(x, -1)
        pd = _bf(p * _drop_mask(p.shape, pdrop, seed)) if pdrop > 0 else None
        return _bf(p), pd

    def softmax_bwd(self, p, dp, T, pdrop=0.0, seed=0):
        pf = p.float()
        g = dp.clone()
        g[..., T:] = 0
        g = g * _drop_mask(p.shape, pdrop, seed)
        ds = pf * (g - (pf * g).sum(-1, keepdim=True))
        return _bf(ds)

    # ---- fused attention (plain fp32 math; `keep` lets tests inject the mask the CUDA kernels regenerate)
    @staticmethod
    def _attn_probs(qkv, H, scale, key_keep):
        B, T, D3 = qkv.shape
        D = D3 // 3
        q, k, v = (t.float().reshape(B, T, H, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
        s = (q @ k.transpose(-1, -2)) * scale
        if key_keep is not None:
            s = s.masked_fill(key_keep[:, None, None, :] == 0, -1e9)
        return q, k, v, torch.softmax(s, -1)

    def attn_fwd(self, qkv, H, scale, key_keep=None, pdrop=0.0, seed=0, keep=None):
        B, T, D3 = qkv.shape
        q, k, v, p = self._attn_probs(qkv, H, scale, key_keep)
        m = self._attn_mask(p.shape, pdrop, seed, keep)
        ctx = ((p * m) @ v).transpose(1, 2).reshape(B, T, D3 // 3)
        return _bf(ctx), torch.zeros(B, H, T)

    def attn_bwd(self, qkv, ctx, dctx, lse, H, scale, key_keep=None, pdrop=0.0, seed=0, keep=None, dbias=None):
        out = self._attn_bwd(qkv, ctx, dctx, lse, H, scale, key_keep, pdrop, seed, keep)
        if dbias is not None:  # column sums of dqkv as stored
            dbias += out.float().sum((0, 1))
        return out

    def _attn_bwd(self, qkv, ctx, dctx, lse, H, scale, key_keep=None, pdrop=0.0, seed=0, keep=None):
        B, T, D3 = qkv.shape
        q, k, v, p = self._attn_probs(qkv, H, scale, key_keep)
        m = self._attn_mask(p.shape, pdrop, seed, keep)
        do = dctx.float().reshape(B, T, H, 64).transpose(1, 2)
        dv = (p * m).transpose(-1, -2) @ do
        g = (do @ v.transpose(-1, -2)) * m
        ds = p * (g - (p * g).sum(-1, keepdim=True)) * scale
        dq, dk = ds @ k, ds.transpose(-1, -2) @ q
        return _bf(torch.cat([t.transpose(1, 2).reshape(B, T, D3 // 3) for t in (dq, dk, dv)], -1))

    @staticmethod
    def _attn_mask(shape, pdrop, seed, keep):
        if keep is not None:
            thr = int(pdrop * 4294967296.0)
            return keep.float() * (4294967296.0 / (4294967296.0 - thr))
        return _drop_mask(shape, pdrop, seed)

    def colsum(self, x, out=None):
        return _into(out, x.float().reshape(-1, x.shape[-1]).sum(0))

    # ---- parameter re-layout
    def cast_multi(self, pairs, cache):
        for s, d in pairs:
            d.copy_(s.reshape(d.shape))

    def conv_pack(self, w, s, want_t, out=None):
        Cout, Cin, k = w.shape
        wk = _bf(w.permute(0, 2, 1).reshape(Cout, -1)).contiguous()
        wts = None
        if want_t:
            wts = [_bf(torch.cat([w[:, :, j].t() for j in range(k) if j % s == p], 1)).contiguous() for p in range(s)]
        if out is not None:  # refresh persistent operand buffers in place
            out[0].copy_(wk)
            for d, t in zip(out[1] or [], wts or []):
                d.copy_(t)
            return out
        return wk, wts

    def conv_unpack(self, dwk, Cin, k, out=None):
        return _into(out, dwk.view(dwk.shape[0], k, Cin).permute(0, 2, 1).contiguous())

    @staticmethod
    def _pc_pack(w, groups, transpose):
        D, cg, k = w.shape
        wg = w.view(groups, cg, cg, k)
        if transpose:
            wg = wg.permute(0, 2, 1, 3)
        out = torch.zeros(groups, cg, k, 64)
        out[..., :cg] = wg.permute(0, 1, 3, 2)
        return _bf(out.reshape(D, k * 64)).contiguous()

    def posconv_pack(self, g, v, want_t, out=None):
        D, cg, k = v.shape
        norm2 = (v * v).sum((0, 1))
        w = g.reshape(1, 1, k) * v / norm2.sqrt()
        groups = D // cg
        res = self._pc_pack(w, groups, False), (self._pc_pack(w, groups, True) if want_t else None), norm2
        if out is not None:
            for d, t in zip(out, res):
                if d is not None and t is not None:
                    d.copy_(t)
            return out
        return res

    def posconv_wn_bwd(self, dwp, g, v, norm2, out=None):
        D, cg, k = v.shape
        groups = D // cg
        dW = dwp.view(groups, k, 64, 64)[:, :, :cg, :cg].permute(0, 3, 2, 1).reshape(D, cg, k)
        nrm = norm2.sqrt()
        t = (dW * v).sum((0, 1))
        dg = (t / nrm).reshape(g.shape)
        dv = g.reshape(1, 1, k) / nrm * (dW - v * t / norm2)
        if out is not None:
            return _into(out[0], dv), _into(out[1], dg)
        return dv, dg

    def dropout(self, x, p, seed):
        return (x.float() * _drop_mask(x.shape, p, seed)).to(x.dtype)

    def mul_dgelu(self, dy, g):
        return _bf(dy.float() * g.float())

    def gelu_bwd(self, dy, z):
        return _bf(dy.float() * gelu_grad(z.float()))

    def log_softmax_fwd(self, x):
        return torch.log_softmax(x, -1)

    def log_softmax_bwd(self, dy, y):
        return _bf(dy - y.exp() * dy.sum(-1, keepdim=True))

    # ---- conv0
    def conv0_stats(self, x, w, k, stride, eps):
        z = torch.nn.functional.conv1d(x.double()[:, None, :], w.double()[:, None, :], stride=stride)  # [B,C,L0]
        mean = z.mean(-1)
        var = z.var(-1, unbiased=False)
        return mean.float(), (var + eps).rsqrt().float(), None  # (the CUDA backend also returns its window moments)

    def _conv0_z(self, x, w, stride):
        return torch.nn.functional.conv1d(x[:, None, :], w[:, None, :], stride=stride).transpose(1, 2)  # [B,L0,C]

    def conv0_fwd(self, x, w, gamma, beta, mean, rstd, k, stride):
        z = self._conv0_z(x, w, stride)
        y = (z - mean[:, None, :]) * rstd[:, None, :] * gamma + beta
        return _bf(gelu(y)).contiguous()

    def conv0_bwd(self, x, w, gamma, beta, mean, rstd, mom, k, stride, da, out=None):
        z = self._conv0_z(x, w, stride)
        xh = (z - mean[:, None, :]) * rstd[:, None, :]
        dy = da.float() * gelu_grad(xh * gamma + beta)
        dgamma, dbeta = (dy * xh).sum((0, 1)), dy.sum((0, 1))
        dz = (rstd[:, None, :] * gamma) * (dy - dy.mean(1, keepdim=True) - xh * (dy * xh).mean(1, keepdim=True))
        B, L0, C = dz.shape
        win = x.unfold(1, k, stride)  # [B, L0, k]
        dw = torch.einsum("blc,blk->ck", dz, win)
        if out is not None:
            return _into(out[0], dw), _into(out[1], dgamma), _into(out[2], dbeta)
        return dw, dgamma, dbeta

    # ---- masks / indices / casts
    # a negative index is a padding entry of a worst-case-length index list: gather -> zero row, others skip it
    def rows_gather(self, src, idx, out_dtype):
        ok = idx >= 0
        out = src[idx.clamp_min(0).long()].to(out_dtype)
        out[~ok] = 0
        return out

    def rows_scatter(self, src, idx, n_rows, out_dtype):
        out = torch.zeros(n_rows, src.shape[-1], dtype=out_dtype)
        ok = idx >= 0
        out[idx[ok].long()] = src[ok].to(out_dtype)
        return out

    def rows_set(self, x, idx, vec):
        x[idx[idx >= 0].long()] = vec.to(x.dtype)

    def rows_set_bwd(self, dx, idx, out=None):
        ii = idx[idx >= 0].long()
        dvec = dx[ii].float().sum(0)
        dx[ii] = 0
        return _into(out, dvec)

    def mask_apply(self, x, row_keep=None, chan_zero=None):
        B, T, C = x.shape
        if row_keep is not None:
            x.view(B * T, C)[row_keep.reshape(-1) == 0] = 0
        if chan_zero is not None:
            x.masked_fill_(chan_zero[:, None, :] != 0, 0)

    def cast(self, x, dtype, out=None):
        return _into(out, x.to(dtype))

    def split3(self, x, b_side, out=None):
        hi = _bf(x)
        lo = _bf(x - hi.float())
        return _into(out, torch.cat([hi, lo, hi] if b_side else [hi, hi, lo], 1).contiguous())

    # ---- quantizer / contrastive
    def vq_fwd(self, z, noise, tau, vars2d, G, n_valid=None):
        R = z.shape[0]
        V = z.shape[1] // G
        zz = z.reshape(R * G, V)
        u = (zz + noise) / tau if noise is not None else zz
        kidx = u.argmax(-1)
        Rv = R if n_valid is None else min(int(n_valid), R)  # rows beyond are padding: out of the statistics
        avg = torch.softmax(zz[:Rv * G], -1).sum(0)
        qbar = avg / (Rv * G)
        ppl = torch.exp(-(qbar * torch.log(qbar + 1e-7)).sum())
        g = torch.arange(R * G) % G
        q = vars2d[g * V + kidx].reshape(R, -1)
        return q, _bf(q), kidx.int(), avg, ppl

    def vq_bwd(self, z, noise, tau, G, vd, a_dot, dq, kidx, avg, ppl, dppl, dvars_out=None, n_valid=None):
        Rall = z.shape[0]
        R = Rall if n_valid is None else min(int(n_valid), Rall)
        if R < Rall:  # padding rows: zero gradient, no codebook contribution
            dzv, dvars = self.vq_bwd(z[:R], noise[:R * G] if noise is not None else None, tau, G, vd,
                                     a_dot[:R] if a_dot is not None else None, dq[:R], kidx[:R * G], avg, ppl, dppl, dvars_out)
            dz = torch.zeros(Rall, z.shape[1], dtype=dzv.dtype)
            dz[:R] = dzv
            return dz, dvars
        V = z.shape[1] // G
        N = R * G
        zz = z.reshape(N, V)
        qbar = avg / N
        dqb = -(dppl * ppl / N) * (torch.log(qbar + 1e-7) + qbar / (qbar + 1e-7))
        s = torch.softmax(zz, -1)
        dz = s * (dqb - (s * dqb).sum(-1, keepdim=True))
        if noise is not None:
            p = torch.softmax((zz + noise) / tau, -1)
            a = a_dot.reshape(N, V)
            dz = dz + p * (a - (p * a).sum(-1, keepdim=True)) / tau
        dvars = torch.zeros(G * V, vd)
        g = torch.arange(N) % G
        dvars.index_add_(0, g * V + kidx.long(), dq.reshape(N, vd))
        return _bf(dz.reshape(R, G * V)), _into(dvars_out, dvars)

    # ---- optimizer side: the pointer table is dereferenced through ctypes (CPU tensors)
    @staticmethod
    def _view(ptr, n):
        import ctypes
        import numpy as np
        if ptr == 0:
            return None
        return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_float * n).from_address(int(ptr))))

    def optim_grad_sqnorm(self, table, chunk_tensor, chunk_off, chunk, partials):
        for c in range(chunk_tensor.numel()):
            row = table[int(chunk_tensor[c])]
            n = int(row[4])
            g = self._view(int(row[1]), n)
            off = int(chunk_off[c])
            partials[c] = 0.0 if g is None else float((g[off:off + chunk].double() ** 2).sum())

    def optim_adamw(self, table, chunk_tensor, chunk_off, chunk, partials, max_norm, grad_scale, lr, beta1, beta2, eps,
                    weight_decay, bc1, bc2_sqrt, scale_grads_only, total_norm_out):
        gs = grad_scale
        if partials is not None:
            total = float(partials[:chunk_tensor.numel()].double().sum().sqrt()) * abs(grad_scale)
            if total_norm_out is not None:
                total_norm_out.fill_(total)
            if max_norm > 0:
                gs = grad_scale * min(max_norm / (total + 1e-6), 1.0)
        for row in table:
            n = int(row[4])
            g = self._view(int(row[1]), n)
            if g is None:
                continue
            if scale_grads_only:
                g.mul_(gs)
                continue
            p, m, v = self._view(int(row[0]), n), self._view(int(row[2]), n), self._view(int(row[3]), n)
            gg = g * gs
            p.mul_(1 - lr * weight_decay)
            m.lerp_(gg, 1 - beta1)
            v.mul_(beta2).addcmul_(gg, gg, value=1 - beta2)
            p.addcdiv_(m, v.sqrt() / bc2_sqrt + eps, value=-lr / bc1)
            if int(row[5]):
                import ctypes
                import numpy as np
                raw = np.ctypeslib.as_array((ctypes.c_int16 * n).from_address(int(row[5])))
                torch.from_numpy(raw).view(torch.bfloat16).copy_(p.to(torch.bfloat16))

    def contrastive_fwd(self, x, y, idx, ppl, n_vars, xe_w, div_w, n_valid=None):
        Rall = x.shape[0]
        K = idx.numel() // Rall
        if n_valid is not None and int(n_valid) < Rall:
            R = int(n_valid)
            loss, ce, saved = self.contrastive_fwd(x[:R], y, idx.reshape(Rall, K)[:R], ppl, n_vars, xe_w, div_w)
            return loss, ce, saved + (Rall,)
        R, C = x.shape
        xn = x.norm(dim=-1).clamp_min(1e-8)
        yn = y.norm(dim=-1).clamp_min(1e-8)
        cand = torch.cat([torch.arange(R)[:, None], idx.reshape(R, K).long()], 1)
        cos = torch.einsum("rc,rkc->rk", x, y[cand]) / (xn[:, None] * yn[cand])
        lse = torch.logsumexp(cos, -1)
        prob = torch.exp(cos - lse[:, None])
        ce = (lse - cos[:, 0]).mean()
        loss = xe_w * ce + (div_w * (n_vars - ppl) / n_vars if ppl is not None else 0.0)
        return loss, ce, (xn, yn, cos, prob, cand)

    def contrastive_bwd(self, x, y, idx, saved, dce, n_valid=None):
        if len(saved) == 6:  # padded call: gradients of the valid rows, zeros for the padding
            Rall = saved[5]
            R = saved[2].shape[0]
            dxv, dy = self.contrastive_bwd(x[:R], y, None, saved[:5], dce)
            dx = torch.zeros_like(x)
            dx[:R] = dxv
            return dx, dy
        xn, yn, cos, prob, cand = saved
        R, C = x.shape
        dcos = prob.clone()
        dcos[:, 0] -= 1.0
        dcos = dcos * (dce / R)
        xh = x / xn[:, None]
        yh = y[cand] / yn[cand][..., None]
        dx = (dcos[..., None] * (yh - cos[..., None] * xh[:, None, :])).sum(1) / xn[:, None]
        dyc = dcos[..., None] * (xh[:, None, :] - cos[..., None] * yh) / yn[cand][..., None]
        dy = torch.zeros_like(y)
        dy.index_add_(0, cand.reshape(-1), dyc.reshape(-1, C))
        return dx, dy

    # ---- device-side draws (csrc/draws.cu): the same Philox4x32-10 counters, Floyd subsets and multiply-high range
    # reduction, literally, so that the kernels can be checked bit for bit
    def span_mask_draw(self, seed, seed_dev, B, T, p_start, mask_length, R_max, device):
        rows, mask = emu_span_mask(_eff_seed(seed, seed_dev), B, T, p_start, mask_length, R_max)
        return torch.from_numpy(rows).to(device), torch.from_numpy(mask).to(device)

    def negatives_draw(self, seed, seed_dev, rows, B, K):
        R_max = rows.numel() - 1
        out = emu_negatives(_eff_seed(seed, seed_dev), int(rows[R_max]), B, K, R_max)
        return torch.from_numpy(out).to(rows.device)

    # ---- ctc (emulated with the installed torch op)
    def ctc_greedy(self, lp, in_len, blank):
        """literal restatement of ctc.py:161-162 per utterance"""
        B, T, V = lp.shape
        out = torch.full((B, T), -1, dtype=torch.int32)
        lens = torch.zeros(B, dtype=torch.int32)
        for b in range(B):
            n = T if in_len is None else min(int(in_len[b]), T)
            toks = lp[b, :n].argmax(dim=-1).unique_consecutive() if n > 0 else torch.zeros(0, dtype=torch.long)
            toks = toks[toks != blank]
            out[b, : toks.numel()] = toks.to(torch.int32)
            lens[b] = toks.numel()
        return out.to(lp.device), lens.to(lp.device)

    def ctc_prep(self, targets, pad, eos, target_lengths, input_lengths):
        keep = (targets != pad) & (targets != eos)
        flat = targets[keep].int()
        off = (torch.cumsum(target_lengths, 0) - target_lengths).int()
        il = torch.cat([input_lengths.int(), torch.zeros(1, dtype=torch.int32)])  # + the loss reduction's ticket
        return flat, off, target_lengths.int(), il

    def ctc_forward(self, lp, flat, off, tl, il, max_S, blank, mean, zero_inf, from_logits=False):
        il = il[:-1]
        with torch.enable_grad():
            lpg = lp.detach().clone().requires_grad_(True)
            lpn = torch.log_softmax(lpg, -1) if from_logits else lpg
            nll = torch.nn.functional.ctc_loss(lpn, flat.long(), il.long(), tl.long(), blank=blank, reduction="none",
                                               zero_infinity=False)
        v = torch.where(torch.isinf(nll), torch.zeros_like(nll), nll) if zero_inf else nll
        loss = (v / tl.clamp_min(1)).mean() if mean else v.sum()
        return loss.detach(), nll.detach(), (lpg, nll)

    def ctc_backward(self, lp, flat, off, tl, il, max_S, blank, alpha, nll, grad_out, mean, zero_inf, from_logits=False,
                     batch_major=False):
        lpg, nll_g = alpha
        B = lp.shape[1]
        scale = grad_out.reshape(-1).expand(B).clone()
        if mean:
            scale = scale / (tl.clamp_min(1) * B)
        scale = torch.where(torch.isinf(nll_g.detach()), torch.zeros_like(scale), scale)
        with torch.enable_grad():
            fin = torch.where(torch.isinf(nll_g), torch.zeros_like(nll_g), nll_g)
            (g,) = torch.autograd.grad((fin * scale).sum(), lpg)
        return torch.nan_to_num(g, nan=0.0, posinf=0.0, neginf=0.0)


# ---------------------------------------------------------------------------------------------------------------
# Philox4x32-10 and the draw algorithms of csrc/draws.cu (numpy, uint64 arithmetic)
# ---------------------------------------------------------------------------------------------------------------
def _eff_seed(seed, seed_dev):
    s = int(seed)
    if seed_dev is not None:
        s += int(seed_dev.reshape(-1)[0].item())
    return s & 0xFFFFFFFFFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, seed):
    """vectorised over the counter words (numpy uint64 arrays holding 32-bit values) -> four uint32-valued arrays"""
    import numpy as np
    M = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & M for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        h0, l0, h1, l1 = p0 >> np.uint64(32), p0 & M, p1 >> np.uint64(32), p1 & M
        c0, c1, c2, c3 = h1 ^ c1 ^ k0, l1, h0 ^ c3 ^ k1, l0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M
        k1 = (k1 + np.uint64(0xBB67AE85)) & M
    return c0, c1, c2, c3


_TAG_NUM, _TAG_START, _TAG_DROP, _TAG_NEG = 1, 2, 3, 4


def _draw32(seed, tag, row, i):
    return int(philox4x32_10(i, row, tag, 0, seed)[0])


def _floyd(marks, n, k, seed, tag, row):
    import numpy as np
    if k <= 0:
        return
    js = np.arange(n - k, n, dtype=np.uint64)
    w = philox4x32_10(js, row, tag, 0, seed)[0]
    ts = (w * (js + np.uint64(1))) >> np.uint64(32)
    for j, t in zip(js.tolist(), ts.tolist()):
        if marks[t]:
            marks[j] = 1
        else:
            marks[t] = 1


def emu_span_mask(seed, B, T, p_start, mask_length, R_max):
    import numpy as np
    L = mask_length
    u0 = float(_draw32(seed, _TAG_NUM, 0, 0) >> 8) * (1.0 / 16777216.0)
    num_mask = int(p_start * float(T) / float(L) + u0)
    span = L
    if T - span <= num_mask:
        span = T - num_mask - 1
    n_start = T - span
    num_mask = min(num_mask, n_start)
    m = np.zeros((B, T), dtype=np.uint8)
    for b in range(B):
        s = np.zeros(T, dtype=np.uint8)
        _floyd(s, n_start, num_mask, seed, _TAG_START, b)
        for t in np.flatnonzero(s[:n_start]):
            m[b, t:min(T, t + L)] = 1
    lens = m.sum(1).astype(np.int64)
    keep = int(lens.min())
    rows = np.full(R_max + 1, -1, dtype=np.int32)
    for b in range(B):
        drop = int(lens[b]) - keep
        pos = np.flatnonzero(m[b])
        if drop > 0:
            s = np.zeros(T, dtype=np.uint8)
            _floyd(s, int(lens[b]), drop, seed, _TAG_DROP, b)
            m[b, pos[s[:len(pos)] != 0]] = 0
            pos = np.flatnonzero(m[b])
        rows[b * keep:(b + 1) * keep] = b * T + pos
    rows[R_max] = B * keep
    return rows, m


def emu_negatives(seed, n_valid, B, K, R_max):
    import numpy as np
    total = R_max * K
    groups = (total + 3) // 4
    g = np.arange(groups, dtype=np.uint64)
    w = np.stack(philox4x32_10(g & np.uint64(0xFFFFFFFF), g >> np.uint64(32), _TAG_NEG, 0, seed), 1).reshape(-1)[:total]
    Tm = n_valid // B
    e = np.arange(total, dtype=np.int64)
    r = e // K
    out = np.zeros(total, dtype=np.int64)
    valid = r < n_valid
    if Tm > 1:
        b, t = r // Tm, r % Tm
        n = ((w * np.uint64(Tm - 1)) >> np.uint64(32)).astype(np.int64)
        n = n + (n >= t)
        out = np.where(valid, n + b * Tm, 0)
    else:
        out = np.where(valid, r, 0)
    return out.astype(np.int32)
