"""torchrun --nproc-per-node 2 scripts/dp_debug.py: which gradients of the arena wrapper differ from the mean of the per-rank ones"""
import datetime, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import wav2vec2 as W, ops
from audio8_b200.parallel import DataParallel
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
cfg = dict(d_model=256, num_heads=4, num_layers=3, d_ff=1024, final_dim=128, num_vq_vars=64, num_vq_groups=2)
torch.manual_seed(0)
model = W.create_model(dropout=0.0, dropout_input=0.0, dropout_features=0.0, **cfg).to(dev).train()
loss_fn = W.create_loss(128, 20)
xs = [(torch.randn(3, 32000, generator=torch.Generator().manual_seed(100 + r)) * 0.1).to(dev) for r in range(2)]
def lg(net, x, seed):
    np.random.seed(seed); torch.manual_seed(seed)
    model.zero_grad(set_to_none=True)
    loss_fn(net, x).backward(); torch.cuda.synchronize()
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
loc = [lg(model, xs[r], 7 + r) for r in range(2)]
want = {k: (loc[0][k] + loc[1][k]) / 2 for k in loc[0]}
net = DataParallel(model)
for rep in range(3):
    got = lg(net, xs[rank], 7 + rank)
    a = net._arena
    for k, p in model.named_parameters():
        sc = want[k].abs().max().item() + 1e-12
        e_mean = (got[k] - want[k]).abs().max().item() / sc
        e_loc = (got[k] - loc[rank][k]).abs().max().item() / sc
        if e_mean > 3e-3 and rank == 0:
            off = (p.grad.data_ptr() - a.buf.data_ptr()) // 4
            print(f"rep {rep} {k}: vs mean {e_mean:.3g}, vs own local {e_loc:.3g}; arena offset {off} (early {a.early}, used {a.used}) slot {a.slots.get(ops.grad_key(p))}", flush=True)
if rank == 0:
    print("done", flush=True)
dist.destroy_process_group()
