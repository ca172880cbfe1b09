"""TEST INFRASTRUCTURE ONLY — generate `tests/golden/*.npz` from the UNMODIFIED reference (dev container only).

    python oracle/gen_golden.py

1. imports `/root/reference/audio8/{wav2vec2,ctc}.py` unmodified through `oracle/shim/`;
2. loads deterministic weights (`ref_params.py`) into the reference models with `strict=True`
   (so the key/shape contract is checked against the real classes);
3. runs the reference forward/backward under recorded seeds, and asserts that the restatement
   (`ref_wav2vec2.py`, `ref_ctc.py`) reproduces it — integer artefacts bit-exactly, floats to 1e-5;
4. writes the small fixtures the tests (CPU and GPU box) compare against.
The GPU box has no `/root/reference`; it only ever sees the committed fixtures.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

from load_reference import load_reference  # noqa: E402
import ref_wav2vec2 as R  # noqa: E402
import ref_ctc as RC  # noqa: E402
import ref_params as P  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

TINY_PRE = dict(d_model=128, num_heads=2, num_layers=2, d_ff=256, final_dim=64, num_vq_vars=24, num_vq_groups=2)
TINY_AC = dict(d_model=128, num_heads=2, num_layers=2, d_ff=256)


def gumbel_noise_like_torch(seed, shape):
    """F.gumbel_softmax draws `-empty_like(logits).exponential_().log()` from the global torch RNG."""
    torch.manual_seed(seed)
    return -torch.empty(shape, dtype=torch.float32).exponential_().log()


def close(a, b, tol=2e-5, what=""):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-12
    assert err <= tol * max(1.0, ref), f"{what}: max err {err} (ref scale {ref})"
    return err


def golden_host_integer(w2v):
    """create_mask / Sampler index generation: bit-exact against the reference functions."""
    out = {}
    for name, (B, T, p, ln, seed) in dict(c2=(6, 749, 0.65, 10, 11), c1=(4, 99, 0.5, 10, 12), chan=(4, 256, 0.1, 64, 13),
                                          short=(3, 24, 0.65, 10, 14)).items():
        np.random.seed(seed)
        ref = w2v.create_mask((B, T), p, ln)
        np.random.seed(seed)
        mine = R.create_mask((B, T), p, ln)
        assert (ref == mine).all(), name
        out[f"mask_{name}_cfg"] = np.array([B, T, int(p * 1000), ln, seed])
        out[f"mask_{name}"] = np.packbits(ref, axis=1)
    for name, (B, Tm, K, seed) in dict(c2=(6, 344, 100, 21), small=(2, 17, 10, 22)).items():
        np.random.seed(seed)
        _, ref = w2v.Sampler(K).negatives(torch.zeros(B, Tm, 4))
        np.random.seed(seed)
        mine = R.sample_negative_indices(B, Tm, K)
        assert (ref.numpy() == mine).all(), name
        out[f"neg_{name}_cfg"] = np.array([B, Tm, K, seed])
        out[f"neg_{name}_sum"] = np.array([ref.numpy().sum(), (ref.numpy() * np.arange(ref.numel()).reshape(ref.shape) % 65521).sum()])
        if name == "small":
            out["neg_small"] = ref.numpy()
    np.savez_compressed(os.path.join(OUT, "host_integer.npz"), **out)
    print("host_integer ok")


def golden_pretrain(w2v):
    cfg = dict(TINY_PRE)
    B, L, K, seed = 2, 16000, 10, 5
    sd = P.pretrain_state_dict(seed=1, **{k: v for k, v in cfg.items() if k != "num_heads"})
    torch.manual_seed(0)
    model = w2v.create_model(dropout=0.0, dropout_input=0.0, dropout_features=0.0, **cfg)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    loss_fn = w2v.create_loss(cfg["num_vq_vars"] * cfg["num_vq_groups"], K)
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(3)) * 0.1
    out = {}
    for mode in ("train", "eval"):
        model.train(mode == "train")
        rec = {}
        orig = loss_fn.sample.negatives

        def spy(y, _orig=orig, _rec=rec):
            negs, idx = _orig(y)
            _rec["idx"] = idx.numpy().copy()
            return negs, idx

        loss_fn.sample.negatives = spy
        model.zero_grad()
        np.random.seed(seed)
        torch.manual_seed(seed)
        # reference forward + loss (wav2vec2.py:377-392); also grab the model outputs with identical draws
        loss = loss_fn(model, x)
        loss.backward()
        np.random.seed(seed)
        torch.manual_seed(seed)
        xo, yo, ppl, tmask = model(x)
        loss_fn.sample.negatives = orig
        tmask = tmask.numpy()
        Tm = int(tmask[0].sum())
        # restatement with the same draws
        np.random.seed(seed)
        tm2 = R.create_mask(tmask.shape, 0.65, 10)
        assert (tm2 == tmask).all()
        for _ in range(cfg["num_layers"]):
            np.random.random()
        idx2 = R.sample_negative_indices(B, Tm, K)
        assert (idx2 == rec["idx"]).all(), "negative indices"
        noise = gumbel_noise_like_torch(seed, (B * Tm * cfg["num_vq_groups"], cfg["num_vq_vars"])) if mode == "train" else None
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        st = R.pretrain_loss(sdg, x, tmask, idx2, n_vars=cfg["num_vq_vars"] * cfg["num_vq_groups"],
                             num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], num_groups=cfg["num_vq_groups"],
                             tau=model.quantizer.curr_temperature, gumbel_noise=noise)
        close(st["x"], xo, what="x")
        close(st["y"], yo, what="y")
        close(st["ppl"], ppl, what="ppl")
        close(st["loss"], loss, what="loss")
        st["loss"].backward()
        gn = {}
        for k, p in model.named_parameters():
            g_ref = p.grad
            g_me = sdg[k].grad
            close(g_me, g_ref, tol=2e-4, what="grad " + k)
            gn[k] = g_ref
        pre = mode + "_"
        out[pre + "time_mask"] = np.packbits(tmask, axis=1)
        out[pre + "neg_idx"] = rec["idx"].astype(np.int32)
        out[pre + "loss"] = np.array([loss.item(), st["ce"].item(), ppl.item()])
        out[pre + "vq_idx"] = st["vq_idx"].numpy().astype(np.int32)
        out[pre + "x"] = xo.detach().numpy()[:, ::7, ::5].copy()
        out[pre + "y"] = yo.detach().numpy()[:, ::3, ::5].copy()
        out[pre + "fx"] = st["fx"].detach().numpy()[:, ::6, ::37].copy()
        names = sorted(gn)
        out[pre + "grad_names"] = np.array(names)
        out[pre + "grad_norms"] = np.array([gn[k].norm().item() for k in names])
        out[pre + "grad_heads"] = np.stack([gn[k].reshape(-1)[:4].numpy() for k in names])
        print(f"pretrain[{mode}] loss {loss.item():.6f} ppl {ppl.item():.4f} Tm {Tm}")
    out["cfg"] = np.array([B, L, K, seed, 1, 3])  # B, L, K, draw seed, weight seed, input seed
    np.savez_compressed(os.path.join(OUT, "pretrain_tiny.npz"), **out)


def golden_acoustic(w2v, ctc_mod):
    from eight_mile_compat import Offsets, sequence_mask
    Offsets.GO, Offsets.PAD = 0, 1  # what train.py:22-23 does at import time
    V, B, L, seed = 32, 3, 16000, 9
    cfg = dict(TINY_AC)
    sd = P.acoustic_state_dict(V, seed=2, **{k: v for k, v in cfg.items() if k != "num_heads"})
    torch.manual_seed(0)
    model = w2v.create_acoustic_model(V, dropout=0.0, freeze_fx=False, **cfg)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model.freeze = False
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(4)) * 0.1
    in_len = torch.tensor([16000, 14000, 11000])
    for b in range(B):
        x[b, in_len[b]:] = 0
    tgt_len = torch.tensor([9, 7, 5])
    gt = torch.Generator().manual_seed(6)
    targets = torch.full((B, 12), Offsets.PAD, dtype=torch.long)
    for b in range(B):
        targets[b, : tgt_len[b]] = torch.randint(4, V, (int(tgt_len[b]),), generator=gt)
        targets[b, tgt_len[b]] = Offsets.EOS
    loss_fn = ctc_mod.CTCLoss()
    out = {}
    for mode in ("train", "eval"):
        model.train(mode == "train")
        model.zero_grad()
        np.random.seed(seed)
        pad_mask = sequence_mask(in_len, L)
        lp, fmask = model(x, pad_mask)
        out_len = fmask.sum(-1)
        loss = loss_fn(lp.transpose(1, 0), out_len, targets, tgt_len)
        loss.backward()
        T = lp.shape[1]
        # restatement with identical draws
        tm = cm = None
        if mode == "train":
            np.random.seed(seed)
            tm = R.create_mask((B, T), 0.5, 10)
            cm = R.create_mask((B, cfg["d_model"]), 0.1, 64)
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        lp2, fm2 = R.acoustic_forward(sdg, x, pad_mask, cfg["num_heads"], cfg["num_layers"], tm, cm)
        assert (fm2 == fmask).all()
        close(lp2, lp, what="log_probs")
        loss2 = RC.ctc_loss_reference(lp2.transpose(1, 0), out_len, targets, tgt_len, blank=0, pad=1, eos=2)
        close(loss2, loss, what="ctc loss")
        loss2.backward()
        gn = {}
        for k, p in model.named_parameters():
            if p.grad is None or sdg[k].grad is None:  # e.g. mask_emb in eval mode
                assert (p.grad is None or p.grad.abs().max() == 0) and (sdg[k].grad is None or sdg[k].grad.abs().max() == 0), k
                gn[k] = torch.zeros_like(p)
                continue
            close(sdg[k].grad, p.grad, tol=2e-4, what="grad " + k)
            gn[k] = p.grad
        pre = mode + "_"
        out[pre + "frame_lengths"] = out_len.numpy()
        out[pre + "loss"] = np.array([loss.item()])
        out[pre + "log_probs"] = lp.detach().numpy()[:, ::4, :].copy()
        out[pre + "greedy"] = np.array([";".join(",".join(map(str, s)) for s in RC.greedy_decode(lp.detach(), out_len, 0))])
        if tm is not None:
            out[pre + "time_mask"] = np.packbits(tm, axis=1)
            out[pre + "channel_mask"] = np.packbits(cm, axis=1)
        names = sorted(gn)
        out[pre + "grad_names"] = np.array(names)
        out[pre + "grad_norms"] = np.array([gn[k].norm().item() for k in names])
        print(f"acoustic[{mode}] loss {loss.item():.5f} frames {out_len.tolist()}")
    out["cfg"] = np.array([V, B, L, seed, 2, 4, 6])
    out["in_len"] = in_len.numpy()
    out["tgt_len"] = tgt_len.numpy()
    out["targets"] = targets.numpy()
    np.savez_compressed(os.path.join(OUT, "acoustic_tiny.npz"), **out)


def golden_ctc(ctc_mod):
    """CTCLoss wrapper cases incl. infeasible rows, repeated labels, empty targets; sum and mean reductions."""
    from eight_mile_compat import Offsets
    Offsets.GO, Offsets.PAD = 0, 1
    out = {}
    g = torch.Generator().manual_seed(31)
    cases = dict(small=(12, 3, 6, [4, 2, 0], [12, 9, 5]), rep=(20, 2, 5, [6, 3], [20, 11]), infeasible=(6, 2, 5, [5, 2], [6, 6]),
                 mid=(120, 5, 32, [30, 24, 11, 1, 17], [120, 100, 77, 50, 119]))
    for name, (T, B, V, tl, il) in cases.items():
        lp = torch.randn(T, B, V, generator=g).log_softmax(-1)
        S = max(tl) + 2
        targets = torch.full((B, S), Offsets.PAD, dtype=torch.long)
        for b in range(B):
            hi = 5 if name == "rep" else V  # few symbols -> many repeats
            targets[b, : tl[b]] = torch.randint(3, hi, (tl[b],), generator=g)
            if name == "infeasible" and b == 0:
                targets[b, : tl[b]] = 3  # 5 identical labels need >= 9 frames, only 6 given
            targets[b, tl[b]] = Offsets.EOS
        il_t, tl_t = torch.tensor(il), torch.tensor(tl)
        for red in ("sum", "mean"):
            lpg = lp.clone().requires_grad_(True)
            loss = ctc_mod.CTCLoss(reduction_type=red)(lpg, il_t, targets, tl_t)
            loss.backward()
            mine = RC.ctc_loss_reference(lp, il_t, targets, tl_t, 0, 1, 2, red)
            close(mine, loss, what=name)
            labels = [targets[b, : tl[b]].numpy() for b in range(B)]
            nll, grad = RC.ctc_numpy(lp.double().numpy(), il, labels, 0)
            if red == "sum":
                fin = np.isfinite(nll)
                close(nll[fin].sum(), loss, tol=1e-5, what=name + " numpy nll")
                close(grad, lpg.grad, tol=1e-4, what=name + " numpy grad")
            out[f"{name}_{red}_loss"] = np.array([loss.item()])
            out[f"{name}_{red}_grad"] = lpg.grad.numpy().copy()
        out[f"{name}_lp"] = lp.numpy()
        out[f"{name}_targets"] = targets.numpy()
        out[f"{name}_il"] = il_t.numpy()
        out[f"{name}_tl"] = tl_t.numpy()
        print("ctc", name, "ok")
    np.savez_compressed(os.path.join(OUT, "ctc_cases.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    w2v, ctc_mod = load_reference()
    golden_host_integer(w2v)
    golden_ctc(ctc_mod)
    golden_pretrain(w2v)
    golden_acoustic(w2v, ctc_mod)
    print("fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
