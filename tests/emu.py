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
                        z[off] = v.to(z.dtype)
                    if g.act == _lib.ACT_GELU:
                        v = gelu(v)
                    if g.aux_mode == _lib.AUX_ADD:
                        v = v + aux[off].float()
                    elif g.aux_mode == _lib.AUX_MUL_GELU_GRAD:
                        v = v * gelu_grad(aux[off].float())
                    if g.c_dtype == _lib.OUT_F32_ATOMIC:
                        c[off] += v
                    else:
                        c[off] = v.to(c.dtype)


class EmuBackend:
    name = "emu"

    def gemm(self, g):
        emu_gemm(g)
