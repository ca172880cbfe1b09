"""Host-vs-GPU timeline of one end-to-end step (pinned input -> device, fwd+bwd, loss.item()), to see where the GPU
starves when the host is synchronised every step.  Marks are placed by monkeypatching (no product code changes):
every CUDA-graph replay, the numpy mask / negative-index draws, backward start/end.  For each mark: host time at
which it was reached and GPU time at which the stream reached it (both relative to the step start)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import wav2vec2 as W  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.manual_seed(rank)
np.random.seed(rank)
model = W.create_model().to(dev).train()
net = model
if world > 1:
    import datetime
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True, bucket_cap_mb=128)
    if rank != 0:
        sys.stdout = open(os.devnull, "w")
loss_fn = W.create_loss(640, 100)
x_host = (torch.randn(6, 240000) * 0.1).pin_memory()
params = list(model.parameters())

marks = []
ON = [False]


def mark(name):
    if ON[0]:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        marks.append((name, time.perf_counter(), e))


def wrap(obj, attr, name):
    orig = getattr(obj, attr)

    def f(*a, **k):
        mark(name + " >")
        r = orig(*a, **k)
        mark(name + " <")
        return r
    setattr(obj, attr, f)


wrap(torch.cuda.CUDAGraph, "replay", "graph.replay")
wrap(W, "create_mask", "create_mask")
wrap(W.Sampler, "indices", "neg indices")
wrap(W.Fn.QuantizerFn, "forward", "quantizer fwd")
wrap(W.Fn.ContrastiveFn, "forward", "contrastive fwd")
wrap(W.Fn.ContrastiveFn, "backward", "contrastive bwd")
wrap(W.Fn.QuantizerFn, "backward", "quantizer bwd")


def step():
    mark("step start")
    x = x_host.to(dev, non_blocking=True)
    mark("h2d issued")
    loss = loss_fn(net, x)
    mark("forward enqueued")
    loss.backward()
    mark("backward enqueued")
    for p in params:
        p.grad = None
    mark("grads dropped")
    v = loss.item()
    mark("loss.item() returned")
    return v


import gc


def _gc_cb(phase, info):
    mark(f"gc {phase} gen{info['generation']}")


gc.callbacks.append(_gc_cb)
if os.environ.get("TL_SWITCH"):
    sys.setswitchinterval(float(os.environ["TL_SWITCH"]))
if os.environ.get("TL_GC_OFF"):
    gc.disable()
for _ in range(14):
    step()
torch.cuda.synchronize()
import threading
import traceback

samples = []
main_id = threading.main_thread().ident
stop = [False]


def sampler():
    while not stop[0]:
        fr = sys._current_frames().get(main_id)
        if fr is not None and ON[0]:
            st = traceback.extract_stack(fr)[-4:]
            samples.append((time.perf_counter(), " <- ".join(f"{os.path.basename(f.filename)}:{f.lineno}:{f.name}" for f in reversed(st))))
        time.sleep(0.0005)


if not os.environ.get('TL_NO_SAMPLER'):
    threading.Thread(target=sampler, daemon=True).start()


def mstat():
    m = torch.cuda.memory_stats()
    return m.get("num_device_alloc", 0), m.get("num_device_free", 0), m.get("num_alloc_retries", 0), m.get("reserved_bytes.all.current", 0) >> 20


for rep in range(int(os.environ.get("TL_STEPS", "12"))):
    marks.clear()
    samples.clear()
    m0 = mstat()
    ON[0] = True
    t0 = time.perf_counter()
    step()
    ON[0] = False
    torch.cuda.synchronize()
    m1 = mstat()
    print(f"allocator: cudaMalloc +{m1[0] - m0[0]}, cudaFree +{m1[1] - m0[1]}, retries +{m1[2] - m0[2]}, reserved {m1[3]} MiB")
    prev = None
    for name, t, e in marks:
        if prev is not None and t - prev[1] > 0.002 and not name.startswith(("neg indices <", "loss.item")):
            print(f"  !! host gap {1e3 * (t - prev[1]):.1f} ms between '{prev[0]}' and '{name}':")
            inside = [s_ for s_ in samples if prev[1] <= s_[0] <= t]
            print(f"     {len(inside)} stack samples in the gap")
            for ts, txt in inside[:: max(1, len(inside) // 6)]:
                print(f"     @{1e3 * (ts - marks[0][1]):7.2f} ms {txt}")
        prev = (name, t)
    e0 = marks[0][2]
    print(f"--- step {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms wall")
    for name, t, e in marks:
        print(f"  {name:26s} host {1e3 * (t - marks[0][1]):7.2f} ms   gpu {e0.elapsed_time(e):7.2f} ms")
