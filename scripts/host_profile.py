"""cProfile of the host side of one pre-training step (where does the enqueue time go?)"""
import cProfile
import os
import pstats
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


def step():
    loss = loss_fn(model, x)
    loss.backward()
    for p in model.parameters():
        p.grad = None


for _ in range(4):
    step()
torch.cuda.synchronize()
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"step {i}: enqueue {1e3 * (t1 - t0):.2f} ms, total {1e3 * (t2 - t0):.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
