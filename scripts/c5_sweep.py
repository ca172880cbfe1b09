"""BASELINE configs[4]: CTC loss + conv feature encoder micro-benchmark sweep against the reference's CPU path.

GPU: audio8_b200 kernels through the C ABI, CUDA events, median of 5.  CPU: what the reference executes for the same
call — `torch.nn.functional.ctc_loss` fwd+bwd (ctc.py:197) and the oracle port of the conv feature encoder
(wav2vec2.py:399-456) fwd+bwd — on all host cores, one repetition per shape (bounded).  Writes a markdown table."""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from audio8_b200 import functional as Fn, ops  # noqa: E402
from audio8_b200.ctc import ctc_loss  # noqa: E402
import ref_params as P  # noqa: E402
import ref_wav2vec2 as R  # noqa: E402

dev = "cuda"
torch.set_num_threads(os.cpu_count() or 1)
out = ["# Round 2 — configs[4]: CTC loss and conv feature encoder sweep, B200 kernels vs the reference's CPU path",
       f"\nCPU: {os.cpu_count()} host cores of the GPU box, torch {torch.__version__}.  GB/s = 2*T*B*V*4 bytes (log-probs read once,",
       "gradient written once) / time.\n", "## CTC loss fwd+bwd (V=32, S=T/5)\n",
       "| T | B | GPU us | GPU GB/s | CPU F.ctc_loss ms | speed-up |", "|---:|---:|---:|---:|---:|---:|"]


def gpu_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


V = 32
for T in (50, 250, 750, 1500):
    for B in (8, 64, 256):
        S = max(T // 5, 1)
        g = torch.Generator().manual_seed(T + B)
        lp = torch.randn(T, B, V, generator=g).log_softmax(-1)
        tg = torch.randint(4, V, (B, S), generator=g)
        tl = torch.full((B,), S, dtype=torch.long)
        il = torch.full((B,), T, dtype=torch.long)
        d = lp.to(dev).requires_grad_(True)
        tgd = tg.to(dev)

        def run():
            d.grad = None
            ctc_loss(d, il, tgd, tl, blank=0, pad=1, eos=2).backward()
        ms = gpu_time(run)
        c = lp.clone().requires_grad_(True)
        t0 = time.perf_counter()
        F.ctc_loss(c, tg, il, tl, blank=0, reduction="sum", zero_infinity=True).backward()
        cpu_ms = (time.perf_counter() - t0) * 1e3
        out.append(f"| {T} | {B} | {ms * 1e3:.0f} | {2 * T * B * V * 4 / ms / 1e6:.1f} | {cpu_ms:.1f} | {cpu_ms / ms:.0f}x |")
        print(out[-1], flush=True)

out += ["\n## Conv feature encoder fwd+bwd (7 layers, 512 channels, GroupNorm on layer 0)\n",
        "| B | samples | frames | GPU ms | GPU audio-s/s | CPU ms | CPU audio-s/s | speed-up |", "|---:|---:|---:|---:|---:|---:|---:|---:|"]
sd = P.pretrain_state_dict(seed=0)
spec = [tuple(c) for c in R.CONV_FEATURES[16]]
ws = [sd[f"feature_extractor.conv_layers.{i}.0.weight"].to(dev).requires_grad_(True) for i in range(7)]
gw = sd["feature_extractor.conv_layers.0.2.weight"].to(dev).requires_grad_(True)
gb = sd["feature_extractor.conv_layers.0.2.bias"].to(dev).requires_grad_(True)
for (B, L, cpuB) in ((8, 16080, 8), (8, 32000, 8), (32, 80080, 4), (6, 240000, 1), (16, 480080, 1)):
    x = torch.randn(B, L, device=dev) * 0.1

    def run():
        y = Fn.ConvFeatureFn.apply(x, spec, gw, gb, *ws)
        y.backward(torch.ones_like(y))
    ms = gpu_time(run, reps=3)
    T = R.conv_out_lengths(L, R.CONV_FEATURES[16])[-1]
    sdc = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith("feature_extractor")}
    xc = torch.randn(cpuB, L) * 0.1
    t0 = time.perf_counter()
    yc = R.conv_feature_extractor(sdc, xc)
    yc.backward(torch.ones_like(yc))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    g_rate, c_rate = B * L / 16000 / (ms * 1e-3), cpuB * L / 16000 / (cpu_ms * 1e-3)
    out.append(f"| {B} | {L} | {T} | {ms:.2f} | {g_rate:.0f} | {cpu_ms:.0f} (B={cpuB}) | {c_rate:.1f} | {g_rate / c_rate:.0f}x |")
    print(out[-1], flush=True)
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "c5_sweep.md")
open(path, "w").write("\n".join(out) + "\n")
