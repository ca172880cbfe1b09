import torch
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        t = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
with torch.cuda.graph(g):
    out = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device="cuda")
for seed in (1, 2, 2, 3, 3):
    torch.manual_seed(seed)
    g.replay()
    print(seed, out.item())
import sys
sys.path.insert(0, ".")
from audio8_b200 import wav2vec2 as W
enc = W.AudioTransformerEncoder(2, 128, 0.1, layers=1, d_ff=256).cuda().train()
h = (torch.randn(2, 49, 128, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
for i, seed in enumerate((1, 2, 3, 3, 3, 4, 4)):
    torch.manual_seed(seed)
    o = enc(h)
    print(i, seed, bool(enc._graph.entries), o.float().sum().item(), o.float().abs().sum().item())
