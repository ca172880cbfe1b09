"""torchrun --nproc-per-node 2 scripts/dp_check.py: gradients of audio8_b200.parallel.DataParallel vs
torch DistributedDataParallel on the same per-rank inputs (full-size base model, dropout 0)."""
import datetime
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import wav2vec2 as W  # noqa: E402
from audio8_b200.parallel import DataParallel  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
torch.manual_seed(0)
model = W.create_model(dropout=0.0, dropout_input=0.0, dropout_features=0.0).to(dev).train()
loss_fn = W.create_loss(640, 100)
x = torch.randn(2, 80000, generator=torch.Generator().manual_seed(10 + rank)).to(dev) * 0.1


def grads(net, reps):
    out = None
    for _ in range(reps):
        np.random.seed(3 + rank)
        torch.manual_seed(3 + rank)
        for p in model.parameters():
            p.grad = None
        loss_fn(net, x).backward()
        torch.cuda.synchronize()
        out = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return out


ref = grads(torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]), 3)
net = DataParallel(model)
got = grads(net, 4)  # 4 steps: eager, graph capture, replays
worst = 0.0
for k in ref:
    d = (got[k] - ref[k]).abs().max().item() / (ref[k].abs().max().item() + 1e-12)
    worst = max(worst, d)
    if d > 2e-3:
        print(f"rank {rank} MISMATCH {k}: rel {d:.3g}")
a = net._arena
n_in = sum(1 for p in model.parameters() if p.grad is not None and a.buf.data_ptr() <= p.grad.data_ptr() < a.buf.data_ptr() + 4 * a.buf.numel())
print(f"rank {rank}: worst relative difference {worst:.3g} over {len(ref)} tensors; arena used {a.used * 4 / 1e6:.0f} MB, {n_in} gradients alias it")
dist.destroy_process_group()
