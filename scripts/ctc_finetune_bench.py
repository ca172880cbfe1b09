"""BASELINE configs[2] per GPU: wav2vec2-base CTC fine-tuning step (char vocab 32), B=8 x 15 s ragged, 150-char targets,
frozen feature encoder (train.py's default), dropout 0.1, time + channel masks: fwd+bwd audio-s/s on one GPU."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import wav2vec2 as W  # noqa: E402
from audio8_b200.ctc import CTCLoss, Offsets  # noqa: E402

Offsets.GO, Offsets.PAD = 0, 1  # train.py:22-27
dev = torch.device("cuda")
torch.manual_seed(0)
np.random.seed(0)
B, L, S, V = 8, 240000, 150, 32
model = W.create_acoustic_model(V).to(dev).train()
model.freeze = False
crit = CTCLoss()
g = torch.Generator().manual_seed(1)
x = (torch.randn(B, L, generator=g) * 0.1).to(dev)
in_len = torch.randint(int(0.7 * L), L + 1, (B,), generator=g)
in_len[0] = L
pad_mask = (torch.arange(L)[None, :] < in_len[:, None]).to(dev)
targets = torch.randint(4, V, (B, S), generator=g).to(dev)
tl = torch.full((B,), S, dtype=torch.long)
params = [p for p in model.parameters()]


def step():
    lp, fmask = model(x, pad_mask)
    loss = crit(lp.transpose(1, 0), fmask.sum(-1), targets, tl)
    loss.backward()
    for p in params:
        p.grad = None
    return loss


for _ in range(12):
    step()
torch.cuda.synchronize()
import gc
gc.collect(); gc.freeze()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
t0 = time.perf_counter()
for _ in range(n):
    step().item()
e2e = (time.perf_counter() - t0) / n * 1e3
print(f"CTC fine-tune fwd+bwd, B={B} x {L / 16000:.0f} s: {ms:.2f} ms/step = {B * L / 16000 / (ms * 1e-3):.0f} audio-s/s resident; "
      f"{e2e:.2f} ms/step with loss.item() every step = {B * L / 16000 / (e2e * 1e-3):.0f} audio-s/s")
