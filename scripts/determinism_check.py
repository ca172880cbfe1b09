"""Run every forward kernel of the encoder repeatedly on identical inputs (tiny shapes of the CUDA-graph test, and the
headline shapes) and report which ones are not bit-reproducible."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import gemm_specs as G  # noqa: E402
from audio8_b200 import ops  # noqa: E402
from audio8_b200.ops import ACT_GELU, AUX_ADD  # noqa: E402

be = ops.backend()
dev, bf = "cuda", torch.bfloat16


def r(*shape, dtype=bf, scale=0.5):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


def check(name, fn, n=30):
    ref = [t.clone() for t in fn()]
    bad = 0
    worst = 0.0
    for _ in range(n):
        junk = torch.randn(1 << 20, device=dev)  # perturb allocator / cache state between runs
        out = fn()
        for a, b in zip(ref, out):
            if not torch.equal(a, b):
                bad += 1
                worst = max(worst, (a.float() - b.float()).abs().max().item())
                break
        del junk
    print(f"{name:50s} {'OK' if bad == 0 else f'NON-DETERMINISTIC {bad}/{n} (max diff {worst:.4g})'}", flush=True)


for (B, T, D, H, F_) in [(2, 49, 128, 2, 256), (6, 749, 768, 12, 3072)]:
    print(f"--- B={B} T={T} D={D} H={H}")
    M = B * T
    x = r(B, T, D)
    seed_t = torch.randint(0, 2 ** 62, (1,), device=dev)
    be.set_seed_source(seed_t)
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    pg, pv = torch.rand(1, 1, 128, device=dev) + 0.5, torch.randn(D, D // 16, 128, device=dev) * 0.05
    wp, wpt, n2 = be.posconv_pack(pg.contiguous(), pv.contiguous(), True)
    check("posconv_pack", lambda: be.posconv_pack(pg.contiguous(), pv.contiguous(), True)[:2])
    pb = torch.randn(D, device=dev)

    def posconv():
        s0, z0 = torch.empty_like(x), torch.empty_like(x)
        be.gemm(G.posconv_fwd(x, wp, s0, pb, 16, 128, 63, z_out=z0))
        return s0, z0
    check("posconv_fwd gemm", posconv)
    check("layernorm_fwd + dropout_y", lambda: be.layernorm_fwd(x, g, b, 1e-5, p_y=0.1, seed_y=11)[:1])
    h = r(B, T, D)
    check("layernorm_fwd + residual dropout", lambda: be.layernorm_fwd(x, g, b, 1e-6, h=h, p_h=0.1, seed_h=12)[:3:2])
    wqkv, bqkv = r(3 * D, D, scale=0.05), torch.randn(3 * D, device=dev)

    def qkv():
        o = torch.empty(M, 3 * D, device=dev, dtype=bf)
        be.gemm(G.linear_fwd(x.view(M, D), wqkv, o, bqkv))
        return (o,)
    check("qkv linear", qkv)
    qkv_t = qkv()[0].view(B, T, 3 * D)
    check("attn_fwd (dropout 0.1)", lambda: be.attn_fwd(qkv_t, H, 0.125, None, 0.1, 7))
    check("attn_fwd (no dropout)", lambda: be.attn_fwd(qkv_t, H, 0.125, None, 0.0, 0))
    ctx, lse = be.attn_fwd(qkv_t, H, 0.125, None, 0.1, 7)
    dctx = r(B, T, D)
    check("attn_bwd (dropout 0.1)", lambda: (be.attn_bwd(qkv_t, ctx, dctx, lse, H, 0.125, None, 0.1, 7),))
    w1, b1 = r(F_, D, scale=0.05), torch.randn(F_, device=dev)

    def ffn1():
        o, z = torch.empty(M, F_, device=dev, dtype=bf), torch.empty(M, F_, device=dev, dtype=bf)
        be.gemm(G.linear_fwd(x.view(M, D), w1, o, b1, act=ACT_GELU, z_out=z))
        return o, z
    check("ffn1 linear + gelu", ffn1)
    hid = ffn1()[0]
    w2 = r(D, F_, scale=0.05)

    def ffn2():
        o = torch.empty(M, D, device=dev, dtype=bf)
        be.gemm(G.linear_fwd(hid, w2, o, pb))
        return (o,)
    check("ffn2 linear", ffn2)
    dy = r(M, F_)

    def dgrad():
        o = torch.empty(M, D, device=dev, dtype=bf)
        be.gemm(G.linear_dgrad(dy, w1, o, aux=x.view(M, D), aux_mode=AUX_ADD))
        return (o,)
    check("ffn1 dgrad + add", dgrad)
    be.set_seed_source(None)
