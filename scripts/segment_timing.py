"""GPU time of the step's segments (CUDA events at segment boundaries, no syncs inside the step)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import wav2vec2 as W  # noqa: E402
from audio8_b200 import graphs  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
np.random.seed(0)
model = W.create_model().to(dev).train()
loss_fn = W.create_loss(640, 100)
x = torch.randn(6, 240000, device=dev) * 0.1
marks = []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


def wrap(seg, name):
    orig = seg.run

    def run(*a, **k):
        mark(name + " begin")
        out = orig(*a, **k)
        mark(name + " end")
        return out
    seg.run = run


wrap(model._front_graph, "front fwd")
wrap(model.encoder._graph, "encoder fwd")


def step():
    mark("step begin")
    loss = loss_fn(model, x)
    mark("loss fwd end")
    loss.backward()
    mark("backward end")
    for p in model.parameters():
        p.grad = None


for _ in range(15):
    step()
torch.cuda.synchronize()
acc = {}
N = 10
for it in range(N):
    marks.clear()
    step()
    mark("next")
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks, marks[1:]):
        acc.setdefault(f"{n0} -> {n1}", []).append(e0.elapsed_time(e1))
tot = 0
for k, v in acc.items():
    m = float(np.median(v))
    tot += m
    print(f"{k:40s} {m:7.3f} ms")
print(f"{'sum':40s} {tot:7.3f} ms")
# serialised variant: sync before each step so that host run-ahead cannot hide anything
ts = []
for it in range(N):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("step with a sync before it: median %.3f ms" % float(np.median(ts)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for it in range(20):
    step()
e1.record()
torch.cuda.synchronize()
print("20 steps back to back: %.3f ms/step" % (e0.elapsed_time(e1) / 20))
