"""Where does the host spend a slow step when it is NOT synchronised with the GPU every step?  30 steps enqueued back to
back; a sampler thread records the main thread's Python stack every millisecond; slow steps print their stack histogram."""
import collections
import os
import sys
import threading
import time
import traceback

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
params = list(model.parameters())


def step():
    loss = loss_fn(model, x)
    loss.backward()
    for p in params:
        p.grad = None


for _ in range(14):
    step()
torch.cuda.synchronize()
import gc
gc.collect(); gc.freeze()
if os.environ.get('GC_OFF'): gc.disable()
samples = []
main_id = threading.main_thread().ident
stop = [False]


def sampler():
    while not stop[0]:
        fr = sys._current_frames().get(main_id)
        if fr is not None:
            st = traceback.extract_stack(fr)[-5:]
            samples.append((time.perf_counter(), " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}:{f.name}" for f in reversed(st))))
        time.sleep(0.001)


threading.Thread(target=sampler, daemon=True).start()
marks, mallocs = [], []
for i in range(int(os.environ.get("STEPS", "30"))):
    t0 = time.perf_counter()
    m0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
    step()
    mallocs.append(torch.cuda.memory_stats().get("num_device_alloc", 0) - m0)
    marks.append((t0, time.perf_counter()))
torch.cuda.synchronize()
stop[0] = True
print("cudaMalloc calls per step:", mallocs, "reserved MiB", torch.cuda.memory_reserved() >> 20)
print("host ms per step:", " ".join(f"{1e3 * (b - a):.1f}" for a, b in marks))
for i, (a, b) in enumerate(marks):
    if b - a > 0.012:
        inside = [s for t, s in samples if a <= t <= b]
        print(f"--- step {i}: {1e3 * (b - a):.1f} ms, {len(inside)} samples")
        for s, n in collections.Counter(inside).most_common(6):
            print(f"   {n:4d}  {s}")
