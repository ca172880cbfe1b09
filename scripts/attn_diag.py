"""Forward attention kernel timed with parts of it switched off (A8_ATTN_DIAG builds): where the 37 us go.
Usage: A8_LIB_TAG=<tag> python scripts/attn_diag.py   (the tagged library must have been built with
A8_NVCC_EXTRA=-DA8_ATTN_DIAG=<bits>); prints one line."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio8_b200 import ops

be = ops.backend()
B, T, H = 6, 749, 12
D = 64 * H
torch.manual_seed(0)
qkv = (torch.randn(B, T, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
dctx = torch.randn(B, T, D, device="cuda").to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=30):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for p in (0.0, 0.1):
    ctx, lse = be.attn_fwd(qkv, H, 0.125, None, p, 7)
    f = timeit(lambda: be.attn_fwd(qkv, H, 0.125, None, p, 7))
    b = timeit(lambda: be.attn_bwd(qkv, ctx, dctx, lse, H, 0.125, None, p, 7))
    print(f"tag={os.environ.get('A8_BUILD_TAG', '')} p={p}: fwd {f[0]:.1f} us (min {f[1]:.1f})  bwd {b[0]:.1f} us (min {b[1]:.1f})")
