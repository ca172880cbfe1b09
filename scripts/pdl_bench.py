"""chain of back-to-back launches (GEMM, LN, attention) timed as a whole: eager stream and CUDA graph; run with A8_PDL=0/1"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import gemm_specs as G  # noqa: E402
from audio8_b200 import ops  # noqa: E402

dev = "cuda"
bf = torch.bfloat16
be = ops.backend()
M, D, F_, B, T, H = 4494, 768, 3072, 6, 749, 12


def r(*shape, dtype=bf):
    return (torch.randn(*shape, device=dev) * 0.1).to(dtype)


x, wqkv, qkv, bq = r(M, D), r(3 * D, D), torch.empty(M, 3 * D, device=dev, dtype=bf), r(3 * D, dtype=torch.float32)
wo, a_out, bo = r(D, D), torch.empty(M, D, device=dev, dtype=bf), r(D, dtype=torch.float32)
g, b_ = torch.ones(D, device=dev), torch.zeros(D, device=dev)
s1 = G.linear_fwd(x, wqkv, qkv, bq)
ctx = torch.empty(B, T, D, device=dev, dtype=bf)


def layerish():
    be.gemm(s1)
    c, lse = be.attn_fwd(qkv.view(B, T, 3 * D), H, 0.125, None, 0.0, 0)
    be.gemm(G.linear_fwd(c.view(M, D), wo, a_out, bo))
    be.layernorm_fwd(x, g, b_, 1e-6, h=a_out)


def gemm_chain():
    for _ in range(8):
        be.gemm(s1)


for name, fn in (("gemm x8", gemm_chain), ("qkv+attn+wo+ln", layerish)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 20
    gr = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        for _ in range(5):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"A8_PDL={os.environ.get('A8_PDL', '1')} {name:18s} eager {eager * 1e3:8.1f} us   graph {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us", flush=True)
