"""Micro-benchmark of the tcgen05 GEMM on the shapes of the wav2vec2-base step (CUDA events, L2-cold via rotation).
usage: python scripts/gemm_bench.py [case-substring]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import gemm_specs as G  # noqa: E402
from audio8_b200 import ops  # noqa: E402
from audio8_b200.ops import ACT_GELU, ACT_GELU_DZ, AUX_ADD, AUX_MUL, AUX_MUL_GELU_GRAD, OUT_F32  # noqa: E402

dev = "cuda"
for _k, _e in (("bn", "BN"), ("split", "SPLIT"), ("cluster", "CL")):  # experiments: force gemm_specs._tiling's choice
    if os.environ.get(_e):
        G.FORCE[_k] = int(os.environ[_e])
bf = torch.bfloat16
be = ops.backend()


def r(*shape, dtype=bf):
    return (torch.randn(*shape, device=dev) * 0.1).to(dtype)


M, D, F_, B, T, H = 4494, 768, 3072, 6, 749, 12
Tp = 752
cases = {}
cases["qkv_fwd 4494x2304x768"] = lambda: G.linear_fwd(r(M, D), r(3 * D, D), torch.empty(M, 3 * D, device=dev, dtype=bf), r(3 * D, dtype=torch.float32))
cases["ffn1_fwd+gelu+z 4494x3072x768"] = lambda: G.linear_fwd(r(M, D), r(F_, D), torch.empty(M, F_, device=dev, dtype=bf), r(F_, dtype=torch.float32), act=ACT_GELU_DZ, z_out=torch.empty(M, F_, device=dev, dtype=bf))
cases["ffn2_fwd 4494x768x3072"] = lambda: G.linear_fwd(r(M, F_), r(D, F_), torch.empty(M, D, device=dev, dtype=bf), r(D, dtype=torch.float32))
cases["wo_fwd 4494x768x768"] = lambda: G.linear_fwd(r(M, D), r(D, D), torch.empty(M, D, device=dev, dtype=bf), r(D, dtype=torch.float32))
cases["ffn2_dgrad*gelu' 4494x3072x768"] = lambda: G.linear_dgrad(r(M, D), r(D, F_), torch.empty(M, F_, device=dev, dtype=bf), aux=r(M, F_).to(torch.float16), aux_mode=AUX_MUL)
cases["ffn1_dgrad+add 4494x768x3072"] = lambda: G.linear_dgrad(r(M, F_), r(F_, D), torch.empty(M, D, device=dev, dtype=bf), aux=r(M, D), aux_mode=AUX_ADD)
cases["ffn_wgrad 3072x768x4494"] = lambda: G.linear_wgrad(r(M, F_), r(M, D), torch.zeros(F_, D, device=dev))
cases["qkv_wgrad 2304x768x4494"] = lambda: G.linear_wgrad(r(M, 3 * D), r(M, D), torch.zeros(3 * D, D, device=dev))
cases["wo_wgrad 768x768x4494"] = lambda: G.linear_wgrad(r(M, D), r(M, D), torch.zeros(D, D, device=dev))
cases["qkv_dgrad+add 4494x768x2304"] = lambda: G.linear_dgrad(r(M, 3 * D), r(3 * D, D), torch.empty(M, D, device=dev, dtype=bf), aux=r(M, D), aux_mode=AUX_ADD)
cases["attn_scores"] = lambda: G.attn_scores(r(B, T, 3 * D), torch.empty(B, H, T, Tp, device=dev), H, 0.125)
cases["attn_context"] = lambda: G.attn_context(r(B, H, T, Tp), r(B, T, 3 * D), torch.empty(B, T, D, device=dev, dtype=bf), H)
cases["attn_dk"] = lambda: G.attn_dk(r(B, H, T, Tp), r(B, T, 3 * D), torch.empty(B, T, 3 * D, device=dev, dtype=bf), H, 0.125)
cases["conv1_fwd"] = lambda: G.conv_fwd(r(B, 47999, 512), r(512, 1536), torch.empty(B, 23999, 512, device=dev, dtype=bf), 3, 2, z_out=torch.empty(B, 23999, 512, device=dev, dtype=bf))
cases["conv1_wgrad"] = lambda: G.conv_wgrad(r(B, 23999, 512), r(B, 47999, 512), torch.zeros(512, 1536, device=dev), 3, 2)
cases["conv2_dgrad_p0"] = lambda: G.conv_dgrad(r(B, 11999, 512), r(512, 1024), torch.empty(B, 23999, 512, device=dev, dtype=bf), 3, 2, 0, aux=r(B, 23999, 512).to(torch.float16))
cases["posconv_fwd"] = lambda: G.posconv_fwd(r(B, T, D), r(D, 128 * 64), torch.empty(B, T, D, device=dev, dtype=bf), r(D, dtype=torch.float32), 16, 128, 63, z_out=torch.empty(B, T, D, device=dev, dtype=bf))

# the library incumbent on the same contraction (cuBLAS through torch.matmul, cuDNN through F.conv1d; bf16 in, bf16 out,
# WITHOUT the fused bias / GELU / residual epilogues our kernel carries: a lower bound on what eager PyTorch spends)
import torch.nn.functional as F  # noqa: E402
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
lib_cases = {
    "qkv_fwd 4494x2304x768": lambda a=r(M, D), w=r(3 * D, D): a @ w.t(),
    "ffn1_fwd+gelu+z 4494x3072x768": lambda a=r(M, D), w=r(F_, D): a @ w.t(),
    "ffn2_fwd 4494x768x3072": lambda a=r(M, F_), w=r(D, F_): a @ w.t(),
    "wo_fwd 4494x768x768": lambda a=r(M, D), w=r(D, D): a @ w.t(),
    "ffn2_dgrad*gelu' 4494x3072x768": lambda a=r(M, D), w=r(D, F_): a @ w,
    "ffn1_dgrad+add 4494x768x3072": lambda a=r(M, F_), w=r(F_, D): a @ w,
    "ffn_wgrad 3072x768x4494": lambda a=r(M, F_), b=r(M, D): a.t() @ b,
    "qkv_wgrad 2304x768x4494": lambda a=r(M, 3 * D), b=r(M, D): a.t() @ b,
    "wo_wgrad 768x768x4494": lambda a=r(M, D), b=r(M, D): a.t() @ b,
    "qkv_dgrad+add 4494x768x2304": lambda a=r(M, 3 * D), w=r(3 * D, D): a @ w,
    "conv1_fwd": lambda a=r(B, 512, 47999), w=r(512, 512, 3): F.conv1d(a, w, stride=2),
    "posconv_fwd": lambda a=r(B, D, T + 127), w=r(D, D // 16, 128): F.conv1d(a, w, groups=16),
}


def time_fn(fn, reps):
    for _ in range(3):
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
    return ts[len(ts) // 2], ts[0]


sel = sys.argv[1] if len(sys.argv) > 1 else ""
reps = int(os.environ.get("REPS", "20"))
for name, mk in cases.items():
    if sel and sel not in name:
        continue
    spec = mk()
    full = spec.spec()
    for _ in range(3):
        be.gemm(spec)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()  # evict L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        be.gemm(spec)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    lib = ""
    if name in lib_cases:
        lmed, lmin = time_fn(lib_cases[name], reps)
        lib = f"   | library {lmed * 1e3:7.1f} us {full.flops / lmed / 1e9:7.1f} TFLOP/s  (ours/library time {med / lmed:.2f})"
    print(f"{name:36s} {med * 1e3:8.1f} us   {full.flops / med / 1e9:7.1f} TFLOP/s   (min {ts[0] * 1e3:.1f} us){lib}", flush=True)

# ---- the 48 weight gradients of the 12-layer stack as ONE grouped launch (a8_gemm_group), vs the sum of the 4 x 12 single ones
if not sel or "group" in sel:
    probs = []
    for _ in range(12):
        for (n_, k_) in ((3 * D, D), (D, D), (F_, D), (D, F_)):
            probs.append(G.linear_wgrad_grouped(r(M, n_), r(M, k_), torch.empty(n_, k_, device=dev)))
    fl = sum(b.spec().flops for b in probs)
    for _ in range(2):
        be.gemm_group(probs)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        be.gemm_group(probs)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    med = ts[len(ts) // 2]
    print(f"{'grouped wgrad 12 layers x 4':36s} {med * 1e3:8.1f} us   {fl / med / 1e9:7.1f} TFLOP/s   (min {ts[0] * 1e3:.1f} us)", flush=True)
    del probs

# ---- fused attention (csrc/attn.cu) at the step's shape
if not sel or "attn" in sel:
    qkv = r(B, T, 3 * D)
    dctx = r(B, T, D)
    for pdrop in (0.0, 0.1):
        ctx, lse = be.attn_fwd(qkv, H, 0.125, None, pdrop, 7)
        for name, fn, fl in (("fwd", lambda: be.attn_fwd(qkv, H, 0.125, None, pdrop, 7), 4),
                             ("bwd", lambda: be.attn_bwd(qkv, ctx, dctx, lse, H, 0.125, None, pdrop, 7), 10)):
            for _ in range(3):
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
            med = ts[len(ts) // 2]
            flops = fl * B * H * T * T * 64
            print(f"fused_attn_{name} p={pdrop:<4}                 {med * 1e3:8.1f} us   {flops / med / 1e9:7.1f} TFLOP/s   (min {ts[0] * 1e3:.1f} us)", flush=True)
