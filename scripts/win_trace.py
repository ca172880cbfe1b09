"""clock64 timeline of the tap-window positional-conv kernel (CTA 0, first tile): per tap when the MMA thread saw the
weight tile, when it had issued the tap's MMAs, and when the producer saw the stage free (a8_gemm_set_trace)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio8_b200 import gemm_specs as G, ops, _lib

be = ops.backend()
lib = _lib.load()
dev, bf = "cuda", torch.bfloat16
B, T, D = 6, 749, 768
r = lambda *s, dtype=bf: (torch.randn(*s, device=dev) * 0.1).to(dtype)
spec = G.posconv_fwd(r(B, T, D), r(D, 128 * 64), torch.empty(B, T, D, device=dev, dtype=bf), r(D, dtype=torch.float32), 16, 128, 63,
                     z_out=torch.empty(B, T, D, device=dev, dtype=torch.float16))
for _ in range(3):
    be.gemm(spec)
torch.cuda.synchronize()
tr = torch.zeros(384, dtype=torch.int64, device=dev)
lib.a8_gemm_set_trace(tr.data_ptr())
be.gemm(spec)
torch.cuda.synchronize()
lib.a8_gemm_set_trace(None)
t = tr.cpu()
t0 = int(t[256])  # producer, tap 0
order = [8 * q + rr for rr in range(8) for q in range(16)]  # visiting order: residue-major
iss = [int(t[128 + j]) - t0 for j in order]
saw = [int(t[j]) - t0 for j in order]
print(f"A8_WIN_DEBUG={os.environ.get('A8_WIN_DEBUG', '')}: cycles per tap (median of issue-to-issue) {sorted(b - a for a, b in zip(iss, iss[1:]))[63]}, "
      f"of which issuing the MMAs + commit {sorted(i - s_ for i, s_ in zip(iss, saw))[64]}")
if os.environ.get("A8_WIN_DEBUG"):
    sys.exit(0)
print("visit#  tap  producer_stage_free  mma_saw_weights  mma_issued   (cycles since the producer's first tap)")
prev = None
for n, j in enumerate(order):
    a, b, c = int(t[256 + j]) - t0, int(t[j]) - t0, int(t[128 + j]) - t0
    d = "" if prev is None else f"  +{c - prev}"
    prev = c
    if n < 40 or n % 16 in (0, 1, 15):
        print(f"{n:4d} {j:4d} {a:10d} {b:10d} {c:10d}{d}")
