"""clock64 timeline of the forward attention kernel (A8_ATTN_TRACE build, A8_LIB_TAG=tr): per traced CTA the softmax
thread's waits per unit, the MMA thread's issue points and the producer's stage waits.  Cycles relative to CTA entry."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio8_b200 import ops, _lib

be = ops.backend()
lib = _lib.load()
lib.a8_attn_set_trace.restype = None
lib.a8_attn_set_trace.argtypes = [ctypes.c_void_p]
B, T, H = 6, 749, 12
D = 64 * H
torch.manual_seed(0)
qkv = (torch.randn(B, T, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
for _ in range(3):
    be.attn_fwd(qkv, H, 0.125, None, 0.1, 7)
torch.cuda.synchronize()
tr = torch.zeros(5 * 3 * 64, dtype=torch.int64, device="cuda")
lib.a8_attn_set_trace(tr.data_ptr())
be.attn_fwd(qkv, H, 0.125, None, 0.1, 7)
torch.cuda.synchronize()
lib.a8_attn_set_trace(None)
t = tr.cpu().view(5, 3, 64)
U = 12
for c, name in enumerate(["cta 0", "cta 1", "cta 150", "cta 300", "cta 431"]):
    sm = t[c, 2]
    t0 = int(sm[0])
    if t0 == 0:
        continue
    rel = lambda v: int(v) - t0
    print(f"== {name}: pdl_wait done {rel(sm[1])}, prologue done {rel(sm[2])}, loop end {rel(sm[3 + 2 * U])}, last PV done {rel(sm[4 + 2 * U])}, stored {rel(sm[5 + 2 * U])}, synced {rel(sm[6 + 2 * U])}")
    print("  softmax thread: unit: [arrive at S_FULL wait, S_FULL seen] ->", " ".join(f"{u}:[{rel(sm[3 + 2 * u])},{rel(sm[4 + 2 * u])}]" for u in range(U)))
    mm = t[c, 1]
    print(f"  mma thread: Q_FULL seen {rel(mm[0])};", " ".join(f"{u}:[kv {rel(mm[1 + 3 * u])}, S issued {rel(mm[2 + 3 * u])}, P seen {rel(mm[3 + 3 * u]) if u else '-'}]" for u in range(U)))
    pr = t[c, 0]
    print("  producer: stage free seen", " ".join(str(rel(pr[j])) for j in range(6)))
