"""rel-L2 error of the fused attention kernels (ctx, dQ, dK, dV) against a float64 softmax attention on the same bf16
inputs.  A8_LIB_TAG selects an alternative build, so two builds can be compared on identical data."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import ops
be = ops.backend()
B, H, T = 2, 4, 749
D = H * 64
g = torch.Generator().manual_seed(3)
qkv = (torch.randn(B, T, 3 * D, generator=g)).bfloat16()
qkv[..., :D] *= float(os.environ.get("QSCALE", "1.5"))
dctx = torch.randn(B, T, D, generator=g).bfloat16()
scale = 0.125
ctx, lse = be.attn_fwd(qkv.cuda(), H, scale, None, 0.0, 1)
dqkv = be.attn_bwd(qkv.cuda(), ctx, dctx.cuda(), lse, H, scale, None, 0.0, 1)
x = qkv.double().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
q, k, v = x[0], x[1], x[2]
p = torch.softmax(q @ k.transpose(-1, -2) * scale, -1)
o = (p @ v).permute(0, 2, 1, 3).reshape(B, T, D)
o.backward(dctx.double())
gq = x.grad.permute(1, 3, 0, 2, 4).reshape(B, T, 3 * D)


def rel(a, b):
    a, b = a.double().cpu(), b.double()
    return ((a - b).norm() / b.norm()).item()


print(f"lib tag '{os.environ.get('A8_LIB_TAG', '')}': ctx {rel(ctx, o.detach()):.5f}  (bf16 rounding of the exact result: "
      f"{rel(o.detach().bfloat16(), o.detach()):.5f})  dQ {rel(dqkv[..., :D], gq[..., :D]):.5f}  dK {rel(dqkv[..., D:2*D], gq[..., D:2*D]):.5f}  "
      f"dV {rel(dqkv[..., 2*D:], gq[..., 2*D:]):.5f}  lse max err {(lse.cpu().double() - torch.logsumexp(q.detach() @ k.detach().transpose(-1,-2) * scale, -1) / math.log(2)).abs().max().item():.2e}")
# repeat: a race shows up as run-to-run variation
for rep in range(5):
    c2, l2 = be.attn_fwd(qkv.cuda(), H, scale, None, 0.0, 1)
    d2 = be.attn_bwd(qkv.cuda(), c2, dctx.cuda(), l2, H, scale, None, 0.0, 1)
    print(f"  repeat {rep}: ctx identical {torch.equal(c2, ctx)}, rel {rel(c2, o.detach()):.5f}; dqkv identical {torch.equal(d2, dqkv)}")
