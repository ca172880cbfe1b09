"""Where does the end-to-end step lose time against the device-resident step?  (bench.py: e2e 7.67 ms vs 7.22 ms of GPU.)

The e2e loop reads `loss.item()` every step, so the GPU is idle from the moment the loss is ready until the FIRST kernel
of the next step is enqueued.  This script stamps the host clock inside one e2e step — input feed, draws / uploads,
every graph segment's entry and exit, backward, item() — and prints the median over the steps, then a cProfile of the
same loop.  `A8_DEVICE_DRAWS=1` profiles the device-draw mode.
"""
import cProfile
import os
import pstats
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import graphs  # noqa: E402
from audio8_b200 import wav2vec2 as W  # noqa: E402
from audio8_b200.feed import DeviceFeed  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(1234)
np.random.seed(1234)
B, L = 6, 240000
model = W.create_model().to(dev).train()
W.set_prefetch_draws(True)
loss_fn = W.create_loss(640, 100)
params = list(model.parameters())
host_in = (torch.randn(B, L) * 0.1).pin_memory()


def batches():
    while True:
        yield (host_in,)


feed = DeviceFeed(batches(), dev, depth=2)
stamps = []
_orig_run = graphs.GraphedSegment.run


def _run(self, fn, inputs, params_, extra=(), clone_outputs=False):
    stamps.append(("> " + self.name[:24], time.perf_counter()))
    out = _orig_run(self, fn, inputs, params_, extra, clone_outputs)
    stamps.append(("< " + self.name[:24], time.perf_counter()))
    return out


graphs.GraphedSegment.run = _run


def step():
    stamps.append(("step start", time.perf_counter()))
    x = next(feed)[0]
    stamps.append(("input fed", time.perf_counter()))
    loss = loss_fn(model, x)
    stamps.append(("forward enqueued", time.perf_counter()))
    loss.backward()
    stamps.append(("backward enqueued", time.perf_counter()))
    for p in params:
        p.grad = None
    v = loss.item()
    stamps.append(("item() returned", time.perf_counter()))
    return v


for _ in range(16):
    step()
import gc
gc.collect()
gc.freeze()
rec = []
for _ in range(20):
    stamps.clear()
    step()
    rec.append(list(stamps))
names = [n for n, _ in rec[0]]
print(f"device draws: {W._DEVICE_DRAWS[0]}   (ms after step start, median of {len(rec)} steps)")
for i, n in enumerate(names):
    ts = [1e3 * (r[i][1] - r[0][1]) for r in rec if len(r) == len(names)]
    print(f"  {statistics.median(ts):8.3f}  {n}")
# GPU time of the same steps, for the gap
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
x_dev = host_in.to(dev)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    loss = loss_fn(model, x_dev)
    loss.backward()
    for p in params:
        p.grad = None
e1.record()
torch.cuda.synchronize()
print(f"device-resident step: {e0.elapsed_time(e1) / 10:.3f} ms")
t0 = time.perf_counter()
for _ in range(20):
    step()
print(f"e2e step: {(time.perf_counter() - t0) * 50:.3f} ms")
graphs.GraphedSegment.run = _orig_run
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    x = next(feed)[0]
    loss = loss_fn(model, x)
    loss.backward()
    for p in params:
        p.grad = None
    loss.item()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
