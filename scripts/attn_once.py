"""one forward + backward launch of the fused attention kernels at the headline shape (for ncu captures)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio8_b200 import ops

be = ops.backend()
B, T, H = int(os.environ.get("B", 6)), 749, 12
D = 64 * H
torch.manual_seed(0)
qkv = (torch.randn(B, T, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
dctx = torch.randn(B, T, D, device="cuda").to(torch.bfloat16)
for _ in range(2):
    ctx, lse = be.attn_fwd(qkv, H, 0.125, None, 0.1, 7)
    be.attn_bwd(qkv, ctx, dctx, lse, H, 0.125, None, 0.1, 7)
torch.cuda.synchronize()
print("ok")
