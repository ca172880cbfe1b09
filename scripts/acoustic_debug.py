import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import model_cases as MC
import ref_ctc, ref_params as P, ref_wav2vec2 as R
from audio8_b200 import wav2vec2 as W
from audio8_b200.ctc import ctc_loss
cfg = dict(d_model=768, num_heads=12, num_layers=int(os.environ.get("NL", "12")))
V, B, L, S = 32, 8, 240000, 150
sd = P.acoustic_state_dict(V, seed=12, **{k: v for k, v in cfg.items() if k != "num_heads"})
model = W.create_acoustic_model(V, dropout=0.0, freeze_fx=True, **cfg)
model.load_state_dict(sd, strict=True)
model = model.cuda(); model.freeze = False; model.train(True)
g = torch.Generator().manual_seed(6)
x = torch.randn(B, L, generator=g) * 0.1
in_len = torch.randint(int(0.6 * L), L + 1, (B,), generator=g); in_len[0] = L
for b in range(B): x[b, in_len[b]:] = 0
pad_mask = torch.arange(L)[None, :] < in_len[:, None]
tgt_len = torch.randint(S // 2, S + 1, (B,), generator=g)
targets = torch.full((B, S), 1, dtype=torch.long)
for b in range(B): targets[b, : tgt_len[b]] = torch.randint(4, V, (int(tgt_len[b]),), generator=g)
np.random.seed(4)
lp, fmask = model(x.cuda(), pad_mask.cuda())
out_len = fmask.sum(-1)
lp.retain_grad()
loss = ctc_loss(lp.transpose(1, 0), out_len, targets.cuda(), tgt_len, blank=0, pad=1, eos=2)
loss.backward()
T = lp.shape[1]
np.random.seed(4)
tm = R.create_mask((B, T), 0.5, 10); cm = R.create_mask((B, 768), 0.1, 64)
sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
lp2, fm2 = R.acoustic_forward(sdg, x, pad_mask, 12, cfg["num_layers"], tm, cm)
lp2.retain_grad()
loss2 = ref_ctc.ctc_loss_reference(lp2.double().transpose(1, 0), out_len.cpu(), targets, tgt_len, 0, 1, 2)
loss2.backward()
print("loss ours", loss.item(), "oracle(f64 ctc)", loss2.item())
valid = fm2[..., None].expand_as(lp2)
d = (lp.detach().float().cpu() - lp2.detach())[valid]
print("log-prob err on valid frames: max", d.abs().max().item(), "rms", d.pow(2).mean().sqrt().item(), "ref rms", lp2.detach()[valid].pow(2).mean().sqrt().item())
g1, g2 = lp.grad.float().cpu(), lp2.grad.float()
def cmp(a, b, name):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    print(f"{name}: cos {(a @ b / (a.norm() * b.norm())).item():.5f} rel {((a - b).norm() / b.norm()).item():.4f} |ref| {b.norm().item():.4g}")
cmp(g1, g2, "dL/dlogprob (ours vs oracle)")
# isolate the CTC kernel: our CTC on the ORACLE's log-probs
lp3 = lp2.detach().cuda().requires_grad_(True)
l3 = ctc_loss(lp3.transpose(1, 0), out_len, targets.cuda(), tgt_len, blank=0, pad=1, eos=2)
l3.backward()
print("our CTC on oracle log-probs: loss", l3.item())
cmp(lp3.grad.float().cpu(), g2, "CTC kernel grad vs f64 oracle (same log-probs)")
lp4 = lp2.detach().clone().requires_grad_(True)
l4 = ref_ctc.ctc_loss_reference(lp4.transpose(1, 0), out_len.cpu(), targets, tgt_len, 0, 1, 2); l4.backward()
cmp(lp4.grad, g2, "ATen fp32 CTC grad vs f64 oracle (same log-probs)")
got = dict(model.named_parameters())
for k in ("proj.weight", "proj.bias", "encoder.encoder.transformer.encoders.%d.ffn.3.layer.weight" % (cfg["num_layers"] - 1), "encoder.encoder.transformer.encoders.0.ffn.3.layer.weight", "encoder.mask_emb"):
    cmp(got[k].grad.cpu(), sdg[k].grad, "grad " + k)
