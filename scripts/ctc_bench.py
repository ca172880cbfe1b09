"""CTC loss kernels alone (BASELINE configs[4] sweep corners): forward (alpha/beta recursion) and backward (gradient)
timed separately with CUDA events; GB/s on the algorithmic bytes 2*T*B*V*4."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import ops  # noqa: E402

be = ops.backend()
dev = "cuda"
go = torch.ones((), device=dev)
for (T, B, S, V) in [(50, 8, 10, 32), (250, 32, 50, 32), (749, 8, 150, 32), (750, 64, 150, 32), (1500, 256, 300, 32)]:
    lp = torch.randn(T, B, V, device=dev).log_softmax(-1)
    tg = torch.randint(4, V, (B, S), device=dev)
    tl = torch.full((B,), S, dtype=torch.int64, device=dev)
    il = torch.full((B,), T, dtype=torch.int64, device=dev)
    flat, off, tl32, il32 = be.ctc_prep(tg, 1, 2, tl, il)
    res = {}
    for name in ("fwd", "bwd"):
        ts = []
        for it in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if name == "fwd":
                e0.record()
                out = be.ctc_forward(lp, flat, off, tl32, il32, S, 0, False, True)
                e1.record()
            else:
                l_, n_, a_ = out
                e0.record()
                be.ctc_backward(lp, flat, off, tl32, il32, S, 0, a_, n_, go, False, True)
                e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[name] = sorted(ts)[len(ts) // 2] * 1e3
    tot = res["fwd"] + res["bwd"]
    print(f"T={T:5d} B={B:4d} S={S:4d}: recursion {res['fwd']:8.1f} us  gradient {res['bwd']:7.1f} us  "
          f"{2 * T * B * V * 4 / tot / 1e3:7.2f} GB/s  ({res['fwd'] * 1e3 / T:.0f} ns per time step)", flush=True)
