"""which steps stall on the host, and do they coincide with cudaMalloc / cudaHostAlloc calls?"""
import gc
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import wav2vec2 as W  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
np.random.seed(0)
model = W.create_model().to(dev).train()
loss_fn = W.create_loss(640, 100)
x = torch.randn(6, 240000, device=dev) * 0.1
if os.environ.get("NOGC"):
    gc.disable()


def step():
    t = [time.perf_counter()]
    loss = loss_fn(model, x)
    t.append(time.perf_counter())
    loss.backward()
    t.append(time.perf_counter())
    for p in model.parameters():
        p.grad = None
    t.append(time.perf_counter())
    return [1e3 * (b - a) for a, b in zip(t, t[1:])]


for _ in range(5):
    step()
torch.cuda.synchronize()
hs = getattr(torch.cuda, "host_memory_stats", None)
for i in range(40):
    d0 = torch.cuda.memory_stats()["num_device_alloc"]
    h0 = hs()["num_host_alloc"] if hs else 0
    g0 = sum(s["collections"] for s in gc.get_stats())
    ph = step()
    d1 = torch.cuda.memory_stats()["num_device_alloc"]
    h1 = hs()["num_host_alloc"] if hs else 0
    g1 = sum(s["collections"] for s in gc.get_stats())
    print(f"step {i:2d}: fwd {ph[0]:6.1f} bwd {ph[1]:6.1f} zero {ph[2]:5.1f} ms | cudaMalloc +{d1 - d0} hostAlloc +{h1 - h0} gc +{g1 - g0}", flush=True)
torch.cuda.synchronize()
