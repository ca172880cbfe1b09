"""bisect which autograd Function breaks CUDA-graph capture: python scripts/graph_debug.py <case>"""
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200 import functional as Fn  # noqa: E402
from audio8_b200 import wav2vec2 as W  # noqa: E402

case = sys.argv[1]
dev = "cuda"
torch.manual_seed(0)
m = W.create_model(d_model=128, num_heads=2, num_layers=1, d_ff=256, final_dim=64, num_vq_vars=24, dropout=float(os.environ.get("P", "0.0")),
                   dropout_input=0.0, dropout_features=0.0).to(dev).train()
x = torch.randn(2, 16000, device=dev) * 0.1
if case == "ln":
    inp = (torch.randn(2, 49, 512, device=dev).bfloat16().requires_grad_(True),)
    params = (m.layer_norm.weight, m.layer_norm.bias)
    fn = lambda a, g, b: Fn.layer_norm(a, g, b, 1e-5, want_f32=True)
elif case == "linear":
    inp = (torch.randn(2, 49, 512, device=dev).bfloat16().requires_grad_(True),)
    params = tuple(m.proj_to_input.parameters())
    fn = lambda a, *p: m.proj_to_input(a)
elif case == "conv":
    inp = (x,)
    params = tuple(m.feature_extractor.parameters())
    fn = lambda a, *p: m.feature_extractor.forward_channels_last(a)
elif case == "front":
    inp = (x,)
    params = m._front_params()
    fn = m._front
elif case == "lnlin":
    inp = (torch.randn(2, 49, 512, device=dev).bfloat16().requires_grad_(True),)
    params = (m.layer_norm.weight, m.layer_norm.bias, *m.proj_to_input.parameters())

    def fn(a, *p):
        y, yf = Fn.layer_norm(a, m.layer_norm.weight, m.layer_norm.bias, 1e-5, want_f32=True)
        return m.proj_to_input(y), yf
elif case == "enc":
    inp = (torch.randn(2, 49, 128, device=dev).bfloat16().requires_grad_(True),)
    params = tuple(m.encoder.parameters())
    fn = lambda a, *p: m.encoder(a)
try:
    g = torch.cuda.make_graphed_callables(fn, inp + params, num_warmup_iters=2, allow_unused_input=True)
    out = g(*inp, *params)
    outs = out if isinstance(out, tuple) else (out,)
    sum(o.float().sum() for o in outs).backward()
    torch.cuda.synchronize()
    print(case, "OK")
except Exception:
    traceback.print_exc()
    print(case, "FAILED")
