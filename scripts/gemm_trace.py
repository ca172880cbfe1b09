"""clock64() timeline of the first CTAs of one GEMM launch (a8_gemm_set_trace): where do the producer, the MMA issuer
and the epilogue wait?   python scripts/gemm_trace.py [qkv|ffn1|ffn2|conv1]   (env BN / CL / SPLIT force the tiling)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import _lib, gemm_specs as G, ops  # noqa: E402
from audio8_b200.ops import ACT_GELU  # noqa: E402

for _k, _e in (("bn", "BN"), ("split", "SPLIT"), ("cluster", "CL")):
    if _e in os.environ:
        G.FORCE[_k] = int(os.environ[_e])
dev, bf = "cuda", torch.bfloat16
be = ops.backend()
lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "qkv"
M, D, F_ = 4494, 768, 3072


def r(*s, dtype=bf):
    return (torch.randn(*s, device=dev) * 0.1).to(dtype)


if which == "qkv":
    spec = G.linear_fwd(r(M, D), r(3 * D, D), torch.empty(M, 3 * D, device=dev, dtype=bf), r(3 * D, dtype=torch.float32))
elif which == "ffn1":
    spec = G.linear_fwd(r(M, D), r(F_, D), torch.empty(M, F_, device=dev, dtype=bf), r(F_, dtype=torch.float32), act=ACT_GELU,
                        z_out=torch.empty(M, F_, device=dev, dtype=bf))
elif which == "ffn2":
    spec = G.linear_fwd(r(M, F_), r(D, F_), torch.empty(M, D, device=dev, dtype=bf), r(D, dtype=torch.float32))
elif which == "wgrad":  # one weight-gradient problem with the grouped launch's tiling (256 x 256 pairs, no split-K)
    spec = G.linear_wgrad_grouped(r(M, F_), r(M, D), torch.empty(F_, D, device=dev))
elif which == "convw":
    spec = G.conv_wgrad(r(6, 23999, 512), r(6, 47999, 512), torch.zeros(512, 1536, device=dev), 3, 2)
else:
    spec = G.conv_fwd(r(6, 47999, 512), r(512, 1536), torch.empty(6, 23999, 512, device=dev, dtype=bf), 3, 2,
                      z_out=torch.empty(6, 23999, 512, device=dev, dtype=bf))
sp = spec.spec()
print(f"{which}: M={sp.M} N={sp.N} k_blocks={sp.k_blocks} block_n={sp.block_n} cluster={sp.cluster} split={sp.split_k}")
for _ in range(3):
    be.gemm(spec)
torch.cuda.synchronize()
buf = torch.zeros(4 * 3 * 8 * 4, dtype=torch.int64, device=dev)
lib.a8_gemm_set_trace(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if os.environ.get("FLUSH"):
    torch.empty(256 << 20, dtype=torch.uint8, device=dev).zero_()
e0.record()
be.gemm(spec)
e1.record()
torch.cuda.synchronize()
lib.a8_gemm_set_trace(None)
print(f"launch: {1e3 * e0.elapsed_time(e1):.1f} us")
t = buf.cpu().view(4, 3, 8, 4)
names = {0: ["tile begin", "tile loads issued", "-", "-"], 1: ["wait acc-empty", "acc-empty ok", "first k-block landed", "last commit"],
         2: ["wait acc-full", "acc-full ok", "tile drained", "-"]}
w = t[:, 0, 7, :]
w0 = int(w[:, 0].min())
for cta in range(4):
    print(f"CTA {cta} wall-clock ns since the first CTA entered: entry {int(w[cta, 0]) - w0}, set-up done {int(w[cta, 1]) - w0}, "
          f"griddepcontrol.wait passed {int(w[cta, 2]) - w0}, all roles done {int(w[cta, 3]) - w0}")
t[:, 0, 7, :] = 0
for cta in range(2):
    base = min(int(v) for v in t[cta].reshape(-1) if v > 0)
    print(f"CTA {cta} (cycles since its first stamp)")
    for it in range(8):
        if not (t[cta, :, it] > 0).any():
            continue
        for role, rn in ((0, "producer"), (1, "mma"), (2, "epilogue")):
            ev = [f"{names[role][e]}={int(t[cta, role, it, e]) - base}" for e in range(4) if t[cta, role, it, e] > 0]
            if ev:
                print(f"  tile {it} {rn:9s} " + "  ".join(ev))
