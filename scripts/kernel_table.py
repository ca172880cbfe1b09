"""Per-kernel roofline table at the headline shapes (wav2vec2-base pretrain, B=6 x 15 s; CTC sweep corner):
every hot kernel once, L2 flushed between timed launches, CUDA events on the launching stream.

    python scripts/kernel_table.py [--md out.md] [--once]     (--once: one launch per kernel, for `ncu -k regex:`)

`achieved` = ALGORITHMIC bytes (or FLOPs) / time; peaks from MEASURED_PEAKS.json (HBM copy GB/s; bf16 burst TFLOP/s
for a kernel timed alone).  Algorithmic sizes follow SURVEY §8(d) / DESIGN.md §3.
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import gemm_specs as G  # noqa: E402
from audio8_b200 import ops  # noqa: E402
from audio8_b200.ops import ACT_GELU, ACT_GELU_DZ, AUX_ADD, AUX_MUL  # noqa: E402

dev = "cuda"
bf = torch.bfloat16
be = ops.backend()
once = "--once" in sys.argv
md_path = sys.argv[sys.argv.index("--md") + 1] if "--md" in sys.argv else None
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    HBM, TF, SRC = PEAKS["hbm_gbs"], PEAKS["bf16_tflops"], "measured"
except Exception:
    HBM, TF, SRC = 6650.0, 1590.0, "fallback"


def r(*shape, dtype=bf, scale=0.1):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


B, L, T, D, F_, H = 6, 240000, 749, 768, 3072, 12
M = B * T
L0 = (L - 10) // 5 + 1
rows = []
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)


def bench(name, fn, bytes_=None, flops=None, note=""):
    reps = 1 if once else 15
    for _ in range(0 if once else 3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    if bytes_ is not None:
        ach, peak, unit, bound = bytes_ / ms / 1e6, HBM, "GB/s", "hbm"
    else:
        ach, peak, unit, bound = flops / ms / 1e9, TF, "TFLOP/s", "tensor"
    rows.append((name, bound, ms * 1e3, ach, unit, ach / peak, note))
    print(f"{name:44s} {ms * 1e3:9.1f} us  {ach:9.1f} {unit:8s} {100 * ach / peak:5.1f}% of {SRC} {bound} peak  {note}", flush=True)


# ---------------------------------------------------------------- conv feature encoder layer 0
x = torch.randn(B, L, device=dev) * 0.1
w0 = (torch.rand(512, 10, device=dev) * 2 - 1) * math.sqrt(0.3)
g0, b0 = torch.ones(512, device=dev), torch.zeros(512, device=dev)
mean, rstd, mom = be.conv0_stats(x, w0, 10, 5, 1e-5)
bench("conv0_stats (moments of x)", lambda: be.conv0_stats(x, w0, 10, 5, 1e-5), bytes_=B * L * 4)
bench("conv0_fwd conv+GroupNorm+GELU", lambda: be.conv0_fwd(x, w0, g0, b0, mean, rstd, 10, 5), bytes_=B * L * 4 + B * L0 * 512 * 2)
da = r(B, L0, 512)
bench("conv0_bwd (single pass)", lambda: be.conv0_bwd(x, w0, g0, b0, mean, rstd, mom, 10, 5, da), bytes_=B * L * 4 + B * L0 * 512 * 2)
del da
# ---------------------------------------------------------------- row kernels
xx, hh = r(M, D), r(M, D)
gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
seed_t = torch.randint(0, 2 ** 62, (1,), device=dev)
be.set_seed_source(seed_t)
y, _, s_, mu_, rs_ = be.layernorm_fwd(xx, gam, bet, 1e-6, h=hh, p_h=0.1, seed_h=5)
bench("layernorm_fwd (+residual, dropout 0.1)", lambda: be.layernorm_fwd(xx, gam, bet, 1e-6, h=hh, p_h=0.1, seed_h=5), bytes_=4 * M * D * 2)
bench("layernorm_bwd (+dropout, dgamma/dbeta/dbias)", lambda: be.layernorm_bwd(xx, s_, mu_, rs_, gam, want_dh=True, p_h=0.1, seed_h=5, want_dbias=True), bytes_=4 * M * D * 2)
big = r(M, F_)
bench("colsum (bias grad) 4494x3072", lambda: be.colsum(big), bytes_=M * F_ * 2)
zz = r(M, F_)
bench("gelu_bwd 4494x3072", lambda: be.gelu_bwd(big, zz), bytes_=3 * M * F_ * 2)
# ---------------------------------------------------------------- quantizer / contrastive (C2: rows 2076)
R, Gq, V, vd, K = 2076, 2, 320, 128, 100
z = torch.randn(R, Gq * V, device=dev) * 3
noise = -torch.empty(R * Gq, V, device=dev).exponential_().log()
vars2d = torch.rand(Gq * V, vd, device=dev)
q, qb, kidx, avg, ppl = be.vq_fwd(z, noise, 0.5, vars2d, Gq)
bench("vq_fwd (gumbel argmax, ppl, codeword gather)", lambda: be.vq_fwd(z, noise, 0.5, vars2d, Gq),
      bytes_=R * Gq * V * 4 * 2 + R * Gq * vd * 6 + Gq * V * vd * 4)
# device-side draws of the step (csrc/draws.cu): span mask of (6, 749) and its 100 negatives per masked step
R_max_d = B * 490
rows_d, _ = be.span_mask_draw(1, None, B, T, 0.65, 10, R_max_d, dev)
bench("span_mask_draw (B=6, T=749: one CTA, Floyd subsets)", lambda: be.span_mask_draw(1, None, B, T, 0.65, 10, R_max_d, dev),
      bytes_=B * T + 4 * R_max_d, note="latency-bound (serial subset draws of one CTA)")
bench("negatives_draw (2940 x 100 indices)", lambda: be.negatives_draw(2, None, rows_d, B, 100), bytes_=R_max_d * 100 * 4)
xc, yc = torch.randn(R, 256, device=dev), torch.randn(R, 256, device=dev)
idx = torch.randint(0, R, (R * K,), device=dev, dtype=torch.int32)
loss, ce, saved = be.contrastive_fwd(xc, yc, idx, ppl, 640.0, 0.1, 10.0)
bench("contrastive_fwd (cos-sim over 1+100, CE)", lambda: be.contrastive_fwd(xc, yc, idx, ppl, 640.0, 0.1, 10.0),
      bytes_=2 * R * 256 * 4 + R * K * 4 + 2 * R * (K + 1) * 4, note="algorithmic bytes incl. saved logits")
dce = torch.ones((), device=dev)
bench("contrastive_bwd", lambda: be.contrastive_bwd(xc, yc, idx, saved, dce), bytes_=4 * R * 256 * 4 + R * K * 4 + 2 * R * (K + 1) * 4)
# ---------------------------------------------------------------- CTC (C5 corner: T=1500, B=256, V=32)
Tc, Bc, Vc, S = 1500, 256, 32, 300
lp = torch.randn(Tc, Bc, Vc, device=dev).log_softmax(-1)
tg = torch.randint(4, Vc, (Bc, S), device=dev)
tl = torch.full((Bc,), S, dtype=torch.int64, device=dev)
il = torch.full((Bc,), Tc, dtype=torch.int64, device=dev)
flat, off, tl32, il32 = be.ctc_prep(tg, 1, 2, tl, il)
lossc, nll, alpha = be.ctc_forward(lp, flat, off, tl32, il32, S, 0, False, True)
go = torch.ones((), device=dev)


def ctc_both():
    l_, n_, a_ = be.ctc_forward(lp, flat, off, tl32, il32, S, 0, False, True)
    be.ctc_backward(lp, flat, off, tl32, il32, S, 0, a_, n_, go, False, True)


bench("ctc fwd+bwd (T=1500,B=256,V=32,S=300)", ctc_both, bytes_=2 * Tc * Bc * Vc * 4, note="latency-bound recursion (T serial steps)")
Tc2, Bc2 = 749, 8
lp2 = torch.randn(Tc2, Bc2, Vc, device=dev).log_softmax(-1)
tg2 = torch.randint(4, Vc, (Bc2, 150), device=dev)
tl2 = torch.full((Bc2,), 150, dtype=torch.int64, device=dev)
il2 = torch.full((Bc2,), Tc2, dtype=torch.int64, device=dev)
f2, o2, t2, i2 = be.ctc_prep(tg2, 1, 2, tl2, il2)


def ctc_small():
    l_, n_, a_ = be.ctc_forward(lp2, f2, o2, t2, i2, 150, 0, False, True)
    be.ctc_backward(lp2, f2, o2, t2, i2, 150, 0, a_, n_, go, False, True)


bench("ctc fwd+bwd (T=749,B=8,V=32,S=150)", ctc_small, bytes_=2 * Tc2 * Bc2 * Vc * 4, note="latency-bound")
# ---------------------------------------------------------------- tensor-core kernels
qkv = r(B, T, 3 * D)
dctx = r(B, T, D)
ctx, lse = be.attn_fwd(qkv, H, 0.125, None, 0.1, 7)
bench("fused attention fwd (dropout 0.1)", lambda: be.attn_fwd(qkv, H, 0.125, None, 0.1, 7), flops=4 * B * H * T * T * 64)
dbq = torch.zeros(3 * D, device=dev)
bench("fused attention bwd (dq + dkv kernels, + QKV bias gradient)", lambda: be.attn_bwd(qkv, ctx, dctx, lse, H, 0.125, None, 0.1, 7, dbias=dbq), flops=8 * B * H * T * T * 64)
gemms = {
    "gemm qkv_fwd 4494x2304x768": lambda: G.linear_fwd(r(M, D), r(3 * D, D), torch.empty(M, 3 * D, device=dev, dtype=bf), r(3 * D, dtype=torch.float32)),
    "gemm ffn1_fwd+gelu 4494x3072x768": lambda: G.linear_fwd(r(M, D), r(F_, D), torch.empty(M, F_, device=dev, dtype=bf), r(F_, dtype=torch.float32), act=ACT_GELU_DZ, z_out=torch.empty(M, F_, device=dev, dtype=bf)),
    "gemm ffn2_fwd 4494x768x3072": lambda: G.linear_fwd(r(M, F_), r(D, F_), torch.empty(M, D, device=dev, dtype=bf), r(D, dtype=torch.float32)),
    "gemm ffn2_dgrad*gelu' 4494x3072x768": lambda: G.linear_dgrad(r(M, D), r(D, F_), torch.empty(M, F_, device=dev, dtype=bf), aux=r(M, F_).to(torch.float16), aux_mode=AUX_MUL),
    "gemm ffn2_dgrad*gelu' + ffn1 bias gradient (colsum epilogue)": lambda: G.linear_dgrad(r(M, D), r(D, F_), torch.empty(M, F_, device=dev, dtype=bf), aux=r(M, F_).to(torch.float16), aux_mode=AUX_MUL, colsum=torch.zeros(F_, device=dev)),
    "gemm ffn_wgrad 3072x768x4494": lambda: G.linear_wgrad(r(M, F_), r(M, D), torch.zeros(F_, D, device=dev)),
    "gemm conv1_fwd (implicit, k=3 s=2)": lambda: G.conv_fwd(r(B, 47999, 512), r(512, 1536), torch.empty(B, 23999, 512, device=dev, dtype=bf), 3, 2, z_out=torch.empty(B, 23999, 512, device=dev, dtype=bf)),
    "gemm conv1_wgrad": lambda: G.conv_wgrad(r(B, 23999, 512), r(B, 47999, 512), torch.zeros(512, 1536, device=dev), 3, 2),
    "gemm conv2_dgrad phase 0 (*gelu')": lambda: G.conv_dgrad(r(B, 11999, 512), r(512, 1024), torch.empty(B, 23999, 512, device=dev, dtype=bf), 3, 2, 0, aux=r(B, 23999, 512).to(torch.float16)),
    "gemm posconv_fwd (k=128, g=16)": lambda: G.posconv_fwd(r(B, T, D), r(D, 128 * 64), torch.empty(B, T, D, device=dev, dtype=bf), r(D, dtype=torch.float32), 16, 128, 63, z_out=torch.empty(B, T, D, device=dev, dtype=bf)),
    "gemm posconv_dgrad (+add)": lambda: G.posconv_dgrad(r(B, T, D), r(D, 128 * 64), torch.empty(B, T, D, device=dev, dtype=bf), 16, 128, 63, aux=r(B, T, D)),
    "gemm posconv_wgrad": lambda: G.posconv_wgrad(r(B, T, D), r(B, T, D), torch.empty(16, 128 * 64, 64, device=dev), 16, 128, 63),
}
for name, mk in gemms.items():
    spec = mk()
    fl = spec.spec().flops
    bench(name, lambda: be.gemm(spec), flops=fl)
# the 48 weight-gradient GEMMs of the 12-layer stack as ONE grouped persistent launch (a8_gemm_group)
probs = []
for _ in range(12):
    for (n_, k_) in ((3 * D, D), (D, D), (F_, D), (D, F_)):
        probs.append(G.linear_wgrad_grouped(r(M, n_), r(M, k_), torch.empty(n_, k_, device=dev)))
bench("gemm grouped wgrad, 12 layers x 4 problems, one launch", lambda: be.gemm_group(probs), flops=sum(b.spec().flops for b in probs))
del probs
# optimizer side: fused clip-grad-norm + AdamW over the base model's 95 M parameters (norm pass 4 B/elem, update 28 B/elem)
from audio8_b200.optim import FusedAdamW  # noqa: E402
ps = [torch.nn.Parameter(torch.randn(n, device=dev) * 0.02) for n in (590 * 1000,) * 160 + (768,) * 40]
for p_ in ps:
    p_.grad = torch.randn_like(p_) * 1e-3
opt = FusedAdamW(ps, lr=2e-4, weight_decay=0.01)
nel = sum(p_.numel() for p_ in ps)
bench(f"fused clip + AdamW step ({nel / 1e6:.0f} M parameters, 2 launches)", lambda: opt.step(clip=1.0), bytes_=nel * 32)

if md_path:
    with open(md_path, "w") as f:
        f.write(f"| kernel | bound | time us | achieved | frac of {SRC} peak | note |\n|---|---|---:|---:|---:|---|\n")
        for (name, bound, us, ach, unit, frac, note) in rows:
            f.write(f"| {name} | {bound} | {us:.1f} | {ach:.0f} {unit} | {frac:.2f} | {note} |\n")
