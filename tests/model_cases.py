"""Shared end-to-end parity cases: the product modules (audio8_b200.wav2vec2 / ctc) against the CPU oracle
(oracle/ref_wav2vec2.py, pinned to the unmodified reference by oracle/gen_golden.py) and the committed golden
fixtures.  Run on CPU through the ABI emulation (host logic) and on the GPU through the CUDA kernels.

Tolerances (bf16 activations / fp32 accumulation vs the fp32 reference; stated once, not tuned per run):
  activations  max|err| <= 3e-2 * max|ref|      loss  rel 1e-2
  gradients    SURVEY §8c's policy: cosine >= 0.999 and rel-L2 <= 3e-2 per parameter tensor; the exceptions are listed
               one by one in `grad_tol` with the measured value and its cause (profiles/r02_parity.md holds every
               tensor's measured cosine / rel-L2, written by the GPU test session itself: see conftest.py)
Integer artefacts (time mask, negative indices, frame lengths, VQ arg-max with shared noise) are bit-exact.
"""
import os

import numpy as np
import torch

import ref_ctc
import ref_params as P
import ref_wav2vec2 as R

GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")
TINY_PRE = dict(d_model=128, num_heads=2, num_layers=2, d_ff=256, final_dim=64, num_vq_vars=24, num_vq_groups=2)
TINY_AC = dict(d_model=128, num_heads=2, num_layers=2, d_ff=256)


def gumbel_noise_like_torch(seed, shape):
    torch.manual_seed(seed)
    return -torch.empty(shape, dtype=torch.float32).exponential_().log()


def act_close(got, want, what, tol=3e-2):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item() + 1e-6
    assert err <= tol * scale, f"{what}: max abs err {err:.4g} vs scale {scale:.4g}"


PARITY_LOG = []  # (case, tensor, cosine, rel-L2, cos_min, rel_max) of every gradient comparison of this session
CASE = ["?"]
SOFT = os.environ.get("A8_PARITY_RECORD") == "1"  # measure-only mode: record every value, fail on none


def _cos_rel(got, want):
    got, want = got.detach().double().cpu().reshape(-1), want.detach().double().cpu().reshape(-1)
    cos = (got @ want / (got.norm() * want.norm() + 1e-30)).item()
    return cos, ((got - want).norm() / (want.norm() + 1e-30)).item()


def grad_close(got, want, what, cos_min=0.999, rel_max=3e-2, floor=None):
    """floor: the same gradient from the ORACLE run with every activation the CUDA path stores in bf16 rounded to bf16
    (values and gradients; all arithmetic fp32: `ref_wav2vec2.set_storage_emulation`).  Its distance to the plain fp32
    oracle is the noise floor of the storage format for this tensor.  A tensor outside (cos_min, rel_max) still passes
    when it is within 1.5x that floor: the deviation is then the storage format's, not the kernels'.  Both are reported."""
    nw = want.detach().double().norm().item()
    if nw < 1e-10:
        assert got.detach().double().norm().item() < 1e-6, f"{what}: expected ~0 gradient"
        return
    cos, rel = _cos_rel(got, want)
    fl = _cos_rel(floor, want) if floor is not None else None
    PARITY_LOG.append((CASE[0], what, cos, rel, cos_min, rel_max, fl))
    if SOFT:
        return
    ok = cos >= cos_min and rel <= rel_max
    if not ok and fl is not None:
        ok = rel <= 1.5 * fl[1] and (1.0 - cos) <= 2.25 * (1.0 - fl[0])
    assert ok, (f"{what}: cosine {cos:.5f}, rel-L2 {rel:.4f} (need {cos_min} / {rel_max}"
                + (f"; bf16-storage oracle floor: cosine {fl[0]:.5f}, rel-L2 {fl[1]:.4f})" if fl else ")"))


def check_param_grads(named_got, want, pinned=True, label="grad "):
    """want: dict name -> reference grad.  The key-projection bias gradient is zero in exact arithmetic (softmax is
    shift invariant along keys); in bf16 it is rounding noise, bounded here against the query-bias gradient."""
    for k, g in named_got:
        g = g if g is not None else torch.zeros_like(want[k])
        if k.endswith("w_K.layer.bias"):
            ref = want[k.replace("w_K", "w_Q")].norm().item()
            assert g.norm().item() <= 0.1 * ref + 1e-6, f"grad {k}: |g| {g.norm().item():.3g} vs |dq bias| {ref:.3g}"
            continue
        grad_close(g, want[k], label + k, **grad_tol(k, pinned))


def upstream_of_quantizer(name):
    return "feature_extractor" in name or name.startswith("layer_norm") or "quantizer.weight_proj" in name


def grad_tol(name, pinned=True):
    """SURVEY §8c's 0.999 / 3e-2 for every tensor when the oracle is evaluated at OUR quantizer logits (`force_z`).
    With free-running logits the tensors upstream of the quantizer inherit the sensitivity of softmax((z + noise) / 0.5)
    at |z| ~ 20 (weight_proj ~ N(0,1), wav2vec2.py:486): a 2^-9 relative (bf16) perturbation of the conv features moves the
    probabilities by ~10 %.  Measured on the B200: cosine 0.9937-0.9990, rel-L2 0.045-0.112 (profiles/r02_parity.md), the
    same value on every conv layer of a case, i.e. one perturbation at the logits carried through linear maps — and the
    same tensors sit at <= 3e-2 once the logits are pinned, which is the check that holds the kernels to account."""
    if not pinned and upstream_of_quantizer(name):
        return dict(cos_min=0.99, rel_max=0.15)
    return {}


def run_pretrain_case(device, mode="train"):
    from audio8_b200 import wav2vec2 as W
    gold = np.load(os.path.join(GOLD_DIR, "pretrain_tiny.npz"))
    B, L, K, seed, wseed, xseed = (int(v) for v in gold["cfg"])
    cfg = dict(TINY_PRE)
    sd = P.pretrain_state_dict(seed=wseed, **{k: v for k, v in cfg.items() if k != "num_heads"})
    model = W.create_model(dropout=0.0, dropout_input=0.0, dropout_features=0.0, **cfg)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model = model.to(device)
    model.train(mode == "train")
    loss_fn = W.create_loss(cfg["num_vq_vars"] * cfg["num_vq_groups"], K)
    x = (torch.randn(B, L, generator=torch.Generator().manual_seed(xseed)) * 0.1)
    tmask = np.unpackbits(gold[mode + "_time_mask"], axis=1)[:, : R.conv_out_lengths(L, R.CONV_FEATURES[16])[-1]].astype(bool)
    Tm = int(tmask[0].sum())
    noise = None
    if mode == "train":
        noise = gumbel_noise_like_torch(seed, (B * Tm * cfg["num_vq_groups"], cfg["num_vq_vars"]))
        model.quantizer.noise_override = noise.to(device)
    model.quantizer.keep_logits = True
    CASE[0] = f"tiny pretrain fixture ({mode}, {device})"
    np.random.seed(seed)
    loss = loss_fn(model, x.to(device))
    loss.backward()
    z_ours = model.quantizer.last_logits.float().cpu()
    # ---- integer artefacts: bit-exact against the fixture written from the unmodified reference
    np.random.seed(seed)
    xo, yo, ppl, tm = model(x.to(device))
    assert (tm.cpu().numpy() == tmask).all(), "time mask differs from the reference's"
    assert (loss_fn.last_neg_idx.astype(np.int32) == gold[mode + "_neg_idx"]).all(), "negative indices differ"
    # ---- oracle with the same draws: first pinned to the fixture, then with our code indices forced (a bf16
    # near-tie flip of one arg-max would otherwise pollute every downstream comparison; flips are bounded below)
    okw = dict(n_vars=cfg["num_vq_vars"] * cfg["num_vq_groups"], num_heads=cfg["num_heads"],
               num_layers=cfg["num_layers"], num_groups=cfg["num_vq_groups"], tau=0.5, gumbel_noise=noise)
    with torch.no_grad():
        st0 = R.pretrain_loss(sd, x, tmask, gold[mode + "_neg_idx"].astype(np.int64), **okw)
    assert abs(st0["loss"].item() - gold[mode + "_loss"][0]) < 1e-4, "oracle drifted from the reference fixture"
    assert (st0["vq_idx"].numpy() == gold[mode + "_vq_idx"]).all()
    kidx = model.quantizer.last_indices.cpu().numpy()
    vq_match = (kidx == gold[mode + "_vq_idx"]).mean()
    assert vq_match >= 0.9, f"VQ arg-max agreement {vq_match:.3f} (bf16 features feed fp32-accurate logits)"
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    st = R.pretrain_loss(sdg, x, tmask, gold[mode + "_neg_idx"].astype(np.int64), force_idx=kidx, **okw)
    st["loss"].backward()
    # second oracle pass evaluated at OUR quantizer logits: every gradient is held to 0.999 / 3e-2 there (see grad_tol)
    sdz = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    R.pretrain_loss(sdz, x, tmask, gold[mode + "_neg_idx"].astype(np.int64), force_idx=kidx, force_z=z_ours, **okw)["loss"].backward()
    check_param_grads([(k, p.grad) for k, p in model.named_parameters()], {k: v.grad for k, v in sdz.items()},
                      pinned=True, label="grad [logits pinned] ")
    act_close(xo, st["x"], "x (context outputs)")
    act_close(xo.cpu()[:, ::7, ::5], torch.from_numpy(gold[mode + "_x"]), "x vs fixture")
    act_close(yo, st["y"], "y (quantized targets)")
    assert abs(ppl.item() - st["ppl"].item()) <= 2e-2 * st["ppl"].item(), (ppl.item(), st["ppl"].item())
    assert abs(loss.item() - st["loss"].item()) <= 1e-2 * abs(st["loss"].item()), (loss.item(), st["loss"].item())
    check_param_grads([(k, p.grad) for k, p in model.named_parameters()], {k: v.grad for k, v in sdg.items()},
                      pinned=False, label="grad [free logits] ")
    return loss.item(), st["loss"].item(), vq_match


def run_acoustic_case(device, mode="train"):
    from audio8_b200 import wav2vec2 as W
    from audio8_b200.ctc import ctc_loss
    gold = np.load(os.path.join(GOLD_DIR, "acoustic_tiny.npz"))
    V, B, L, seed, wseed, xseed, _ = (int(v) for v in gold["cfg"])
    cfg = dict(TINY_AC)
    sd = P.acoustic_state_dict(V, seed=wseed, **{k: v for k, v in cfg.items() if k != "num_heads"})
    model = W.create_acoustic_model(V, dropout=0.0, freeze_fx=False, **cfg)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model = model.to(device)
    model.freeze = False
    model.train(mode == "train")
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(xseed)) * 0.1
    in_len = torch.from_numpy(gold["in_len"])
    for b in range(B):
        x[b, in_len[b]:] = 0
    targets, tgt_len = torch.from_numpy(gold["targets"]), torch.from_numpy(gold["tgt_len"])
    pad_mask = torch.arange(L)[None, :] < in_len[:, None]
    np.random.seed(seed)
    lp, fmask = model(x.to(device), pad_mask.to(device))
    out_len = fmask.sum(-1)
    assert (out_len.cpu().numpy() == gold[mode + "_frame_lengths"]).all(), "frame lengths differ"
    loss = ctc_loss(lp.transpose(1, 0), out_len, targets.to(device), tgt_len, blank=0, pad=1, eos=2)
    loss.backward()
    T = lp.shape[1]
    tm = cm = None
    if mode == "train":
        tm = np.unpackbits(gold[mode + "_time_mask"], axis=1)[:, :T].astype(bool)
        cm = np.unpackbits(gold[mode + "_channel_mask"], axis=1)[:, : cfg["d_model"]].astype(bool)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lp2, fm2 = R.acoustic_forward(sdg, x, pad_mask, cfg["num_heads"], cfg["num_layers"], tm, cm)
    loss2 = ref_ctc.ctc_loss_reference(lp2.transpose(1, 0), out_len.cpu(), targets, tgt_len, 0, 1, 2)
    loss2.backward()
    assert abs(loss2.item() - gold[mode + "_loss"][0]) < 1e-3 * abs(loss2.item()), "oracle drifted from the fixture"
    valid = fm2[..., None].expand_as(lp2)
    act_close(torch.where(valid, lp.detach().float().cpu(), torch.zeros(())), torch.where(valid, lp2.detach(), torch.zeros(())),
              "log-probs (valid frames)", tol=5e-2)
    assert abs(loss.item() - loss2.item()) <= 2e-2 * abs(loss2.item()), (loss.item(), loss2.item())
    check_param_grads([(k, p.grad) for k, p in model.named_parameters()],
                      {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sdg.items()})
    return loss.item(), loss2.item()


def run_graph_case():
    """CUDA-graph replays of the static segments (audio8_b200/graphs.py) against the eager launch sequence of the
    same kernels: same loss / gradients (up to fp32 atomic ordering), and dropout that still varies per replay."""
    from audio8_b200 import graphs
    from audio8_b200 import wav2vec2 as W
    cfg = dict(TINY_PRE)
    torch.manual_seed(0)
    np.random.seed(0)
    model = W.create_model(dropout=0.0, dropout_input=0.0, dropout_features=0.0, **cfg).cuda().train()
    loss_fn = W.create_loss(cfg["num_vq_vars"] * cfg["num_vq_groups"], 10)
    x = torch.randn(2, 16000, device="cuda") * 0.1

    def step():
        np.random.seed(5)
        torch.manual_seed(5)
        model.zero_grad(set_to_none=True)
        loss = loss_fn(model, x)
        loss.backward()
        return loss.item(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

    was = graphs.ENABLED
    try:
        graphs.set_enabled(False)
        l0, g0 = step()
        graphs.set_enabled(True)
        for _ in range(3):
            l1, g1 = step()
        assert model._front_graph.entries and model.encoder._graph.entries, "segments were not captured"
        assert abs(l0 - l1) <= 1e-4 * abs(l0), (l0, l1)
        assert set(g0) == set(g1)
        for k in g0:
            grad_close(g1[k], g0[k], "graph vs eager grad " + k, cos_min=0.9999, rel_max=1e-2)
        # dropout: fresh masks on every replay, reproducible under torch.manual_seed
        enc = W.AudioTransformerEncoder(2, 128, 0.1, layers=1, d_ff=256).cuda().train()
        h = (torch.randn(2, 49, 128, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
        outs = []
        for i in range(5):
            torch.manual_seed(100 + (i if i < 4 else 3))
            outs.append(enc(h).detach().float().clone())
        assert enc._graph.entries
        assert (outs[2] - outs[3]).abs().max().item() > 1e-3, "dropout mask did not change between graph replays"
        d34 = (outs[3] - outs[4]).abs()
        assert torch.equal(outs[3], outs[4]), \
            f"same torch seed must give the same dropout mask: {int((d34 > 0).sum())} of {d34.numel()} differ, max {d34.max().item():.4g}"
    finally:
        graphs.set_enabled(was)
    return l0, l1


FULL_SIZE_GRADS = ("mask_emb", "final_proj.layer.weight", "project_q.layer.weight",
                   "encoder.transformer.encoders.0.ffn.0.layer.weight",
                   "encoder.transformer.encoders.0.self_attn.w_O.layer.weight",
                   "encoder.ln.weight", "proj_to_input.layer.bias")
# what the full-size case adds: the conv feature encoder (32 % of the FLOPs), its GroupNorm, the
# post-extractor LayerNorm, the quantizer and the weight-normed positional conv
FULL_SIZE_GRADS_FRONT = ("feature_extractor.conv_layers.0.0.weight", "feature_extractor.conv_layers.1.0.weight",
                         "feature_extractor.conv_layers.6.0.weight", "feature_extractor.conv_layers.0.2.weight",
                         "layer_norm.weight", "quantizer.weight_proj.weight", "quantizer.vars",
                         "encoder.pos_conv.conv.1.weight_g", "encoder.pos_conv.conv.1.weight_v")


def run_pretrain_generic(device, cfg, B, L, K, seed=3, check_grads=FULL_SIZE_GRADS, split_min=None, train=False,
                         layer_drop=0.0, sample_rate=16, case=None, bf16_floor=False):
    """Any configuration / size (no committed fixture): the product and the oracle are driven from the same numpy seed
    (the oracle's create_mask / sample_negative_indices are pinned to the reference by test_oracle.py), dropout 0.
    train=False: eval-mode quantizer (arg-max, no Gumbel noise); train=True: training mode with shared Gumbel noise and,
    with layer_drop > 0, LayerDrop decided by the same numpy draws on both sides (eight_mile: one draw per layer, a layer
    runs iff draw >= layer_drop).  Integer artefacts bit-exact, loss rel 1e-2, the listed gradients ("all" = every
    parameter) by cosine / rel-L2."""
    from audio8_b200 import wav2vec2 as W
    CASE[0] = case or f"pretrain d={cfg.get('d_model', 768)} L={cfg.get('num_layers', 12)} B={B} x {L} K={K}"
    sd = P.pretrain_state_dict(seed=11, sample_rate=sample_rate, **{k: v for k, v in cfg.items() if k != "num_heads"})
    model = W.create_model(sample_rate=sample_rate, dropout=0.0, dropout_input=0.0, dropout_features=0.0,
                           layer_drop=layer_drop, **cfg)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model = model.to(device).train(train)  # eval: arg-max quantizer, masking still applied (reference :937)
    if split_min is not None:
        model.encoder.split_min_layers = split_min  # force the two-segment encoder on shallow test models
    G_, V_ = cfg.get("num_vq_groups", 2), cfg.get("num_vq_vars", 320)
    n_vars = V_ * G_
    n_layers = cfg.get("num_layers", 12)
    loss_fn = W.create_loss(n_vars, K)
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(5)) * 0.1
    # the oracle's draws from the seed, in the reference's order: mask, one draw per layer, negatives
    cf = R.CONV_FEATURES[sample_rate]
    T = R.conv_out_lengths(L, cf)[-1]
    np.random.seed(seed)
    tmask = R.create_mask((B, T), 0.65, 10)
    draws = [np.random.random() for _ in range(n_layers)]
    active = [(not train) or d >= layer_drop for d in draws]
    Tm = int(tmask[0].sum())
    neg = R.sample_negative_indices(B, Tm, K)
    noise = None
    if train:
        noise = gumbel_noise_like_torch(seed, (B * Tm * G_, V_))
        model.quantizer.noise_override = noise.to(device)
        if layer_drop > 0:
            assert not all(active) and any(active), "pick a seed that drops some but not all layers"
    model.quantizer.keep_logits = True
    np.random.seed(seed)
    loss = loss_fn(model, x.to(device))
    loss.backward()
    assert (loss_fn.last_neg_idx.astype(np.int64) == neg).all(), "negative indices differ from the oracle's draws"
    kidx = model.quantizer.last_indices.cpu().numpy()
    z_ours = model.quantizer.last_logits.float().cpu()
    okw = dict(n_vars=n_vars, num_heads=cfg.get("num_heads", 12), num_layers=n_layers, num_groups=G_, tau=0.5,
               gumbel_noise=noise, conv_features=cf, active_layers=active)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    st = R.pretrain_loss(sdg, x, tmask, neg, force_idx=kidx, **okw)
    st["loss"].backward()
    with torch.no_grad():
        st_free = R.pretrain_loss(sd, x, tmask, neg, **okw)
    vq_match = (kidx == st_free["vq_idx"].numpy()).mean()
    VQ_LOG.append((CASE[0], int((kidx != st_free["vq_idx"].numpy()).sum()), int(kidx.size)))
    assert vq_match >= 0.95, f"VQ arg-max agreement {vq_match:.4f}"
    assert abs(loss.item() - st["loss"].item()) <= 1e-2 * abs(st["loss"].item()), (loss.item(), st["loss"].item())
    got = dict(model.named_parameters())
    names = list(got) if check_grads == "all" else list(check_grads)
    # the oracle evaluated at OUR quantizer logits: the pass that holds every gradient to 0.999 / 3e-2 (see grad_tol)
    sdz = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    R.pretrain_loss(sdz, x, tmask, neg, force_idx=kidx, force_z=z_ours, **okw)["loss"].backward()
    floor = {}
    if (bf16_floor or device == "cuda") and torch.cuda.is_available():
        # noise floor of bf16 activation STORAGE: the same oracle, fp32 arithmetic (run on the GPU with TF32 off, for
        # speed), every activation the CUDA path stores in bf16 rounded to bf16 in value and gradient; same draws, same
        # pinned logits; compared against the fp32 CPU oracle
        sdb = {k: v.clone().cuda().requires_grad_(True) for k, v in sd.items()}
        okb = dict(okw, gumbel_noise=noise.cuda() if noise is not None else None)
        tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
        R.set_storage_emulation(True)
        try:
            lb = R.pretrain_loss(sdb, x.cuda(), tmask, neg, force_idx=kidx, force_z=z_ours.cuda(), **okb)["loss"]
            lb.backward()
        finally:
            R.set_storage_emulation(False)
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        floor = {k: v.grad.float().cpu() for k, v in sdb.items() if v.grad is not None}
    for k in names:
        layer = int(k.split("encoders.")[1].split(".")[0]) if "encoders." in k else None
        if layer is not None and not active[layer]:
            assert got[k].grad is None or got[k].grad.abs().max().item() == 0, f"dropped layer got a gradient: {k}"
            continue
        if k.endswith("w_K.layer.bias"):  # zero in exact arithmetic (softmax is shift invariant along keys)
            ref = sdg[k.replace("w_K", "w_Q")].grad.norm().item()
            assert got[k].grad.norm().item() <= 0.1 * ref + 1e-6
            continue
        grad_close(got[k].grad, sdz[k].grad, "grad [logits pinned] " + k, floor=floor.get(k), **grad_tol(k, True))
        if upstream_of_quantizer(k):  # reported (and loosely bounded) with free-running logits as well
            grad_close(got[k].grad, sdg[k].grad, "grad [free logits] " + k, **grad_tol(k, False))
    return loss.item(), st["loss"].item(), vq_match


def check_span_mask(rows, mask, B, T, p_start, mask_length, R_max):
    """structural invariants of a device-drawn span mask (reference wav2vec2.py:189-216)"""
    rows = np.asarray(rows)
    mask = np.asarray(mask).astype(bool).reshape(B, T)
    n = int(rows[R_max])
    assert rows.shape == (R_max + 1,) and n % B == 0 and 0 <= n <= R_max
    assert (rows[n:R_max] == -1).all(), "padding entries must be -1"
    assert (np.flatnonzero(mask.reshape(-1)) == rows[:n]).all(), "row list != nonzero(mask) in row-major order"
    per_row = mask.sum(1)
    assert (per_row == n // B).all(), f"rows hold {per_row} masked frames, expected {n // B} each"
    nm_hi = int(p_start * T / float(mask_length) + 1.0)
    assert n // B <= min(T, nm_hi * mask_length)
    assert n // B >= min(T, (nm_hi - 1)) , "fewer masked frames than non-overlapping single-frame spans would give"
    return n // B


def check_negatives(neg, n_valid, B, K, R_max):
    """never the positive, always inside the utterance (reference wav2vec2.py:955-976); zeros on the padding rows"""
    neg = np.asarray(neg).reshape(R_max, K).astype(np.int64)
    Tm = n_valid // B
    r = np.arange(n_valid)[:, None]
    v = neg[:n_valid]
    assert (v != r).all(), "a negative equals its positive"
    assert (v // Tm == r // Tm).all(), "a negative leaves its utterance"
    assert (neg[n_valid:] == 0).all()


def run_pretrain_device_draws(device, cfg, B, L, K, seed=3, check_grads=("mask_emb", "project_q.layer.weight",
                                                                       "final_proj.layer.weight",
                                                                       "feature_extractor.conv_layers.0.0.weight")):
    """SURVEY 8f-1 device mode: span mask and negatives drawn by csrc/draws.cu inside the step.  The draws are read back
    and handed to the oracle, so loss and gradients are held to the same bars as in the host-draw mode; the draws
    themselves are checked structurally here and bit for bit against the emulation in test_kernels.py."""
    from audio8_b200 import wav2vec2 as W
    CASE[0] = f"pretrain, device draws d={cfg.get('d_model', 768)} B={B} x {L} K={K}"
    sd = P.pretrain_state_dict(seed=11, **{k: v for k, v in cfg.items() if k != "num_heads"})
    model = W.create_model(dropout=0.0, dropout_input=0.0, dropout_features=0.0, **cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(device).eval()
    G_, V_ = cfg.get("num_vq_groups", 2), cfg.get("num_vq_vars", 320)
    loss_fn = W.create_loss(V_ * G_, K)
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(5)) * 0.1
    T = R.conv_out_lengths(L, R.CONV_FEATURES[16])[-1]
    model.quantizer.keep_logits = True
    W.set_device_draws(True)
    try:
        torch.manual_seed(seed)
        state = np.random.get_state()
        with torch.no_grad():
            out = model(x.to(device))
        assert out[1].shape[:2] == (B, model.max_masked_rows(B, T) // B) and out[3].dtype == torch.bool
        loss = loss_fn(model, x.to(device))
        loss.backward()
        assert _same_rng(state, np.random.get_state(), n_layer_draws=2 * cfg.get("num_layers", 12)), \
            "device mode must not draw masks / negatives from numpy"
    finally:
        W.set_device_draws(False)
    R_max = model.max_masked_rows(B, T)
    rows = loss_fn.last_rows.cpu().numpy()
    n = int(rows[R_max])
    tmask = np.zeros(B * T, dtype=bool)
    tmask[rows[:n]] = True
    tmask = tmask.reshape(B, T)
    Tm = check_span_mask(rows, tmask, B, T, 0.65, 10, R_max)
    negp = loss_fn.last_neg_idx.cpu().numpy()
    check_negatives(negp, n, B, K, R_max)
    neg = negp.reshape(R_max, K)[:n].reshape(B, Tm * K).astype(np.int64)
    kidx = model.quantizer.last_indices.cpu().numpy()[:n * G_]
    z_ours = model.quantizer.last_logits.float().cpu()[:n]
    okw = dict(n_vars=V_ * G_, num_heads=cfg.get("num_heads", 12), num_layers=cfg.get("num_layers", 12), num_groups=G_,
               tau=0.5, gumbel_noise=None, conv_features=R.CONV_FEATURES[16])
    sdz = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    st = R.pretrain_loss(sdz, x, tmask, neg, force_idx=kidx, force_z=z_ours, **okw)
    st["loss"].backward()
    assert abs(loss.item() - st["loss"].item()) <= 1e-2 * abs(st["loss"].item()), (loss.item(), st["loss"].item())
    got = dict(model.named_parameters())
    for k in check_grads:
        grad_close(got[k].grad, sdz[k].grad, "grad [device draws] " + k, **grad_tol(k, True))
    return loss.item(), st["loss"].item()


def run_device_draws_replay_case():
    """device draws inside CUDA-graph replays: a fresh mask / fresh negatives on every replay, reproducible under
    torch.manual_seed, structurally valid every time, and the host uploads nothing"""
    from audio8_b200 import graphs
    from audio8_b200 import wav2vec2 as W
    cfg = dict(TINY_PRE)
    B, L, K = 3, 24000, 10
    torch.manual_seed(0)
    model = W.create_model(**cfg).cuda().train()
    loss_fn = W.create_loss(cfg["num_vq_vars"] * cfg["num_vq_groups"], K)
    x = torch.randn(B, L, device="cuda") * 0.1
    T = R.conv_out_lengths(L, R.CONV_FEATURES[16])[-1]
    R_max = model.max_masked_rows(B, T)
    was = graphs.ENABLED
    W.set_device_draws(True)
    try:
        graphs.set_enabled(True)
        seen = []
        for i, seed in enumerate([1, 2, 3, 4, 4, 5]):
            torch.manual_seed(seed)
            model.zero_grad(set_to_none=True)
            loss = loss_fn(model, x)
            loss.backward()
            assert np.isfinite(loss.item())
            rows = loss_fn.last_rows.cpu().numpy().copy()
            neg = loss_fn.last_neg_idx.cpu().numpy().copy()
            n = int(rows[R_max])
            tmask = np.zeros(B * T, dtype=bool)
            tmask[rows[:n]] = True
            check_span_mask(rows, tmask, B, T, 0.65, 10, R_max)
            check_negatives(neg, n, B, K, R_max)
            assert model.mask_emb.grad is not None and model.mask_emb.grad.abs().sum().item() > 0
            seen.append((rows, neg, loss.item()))
        assert model._front_graph.entries and loss_fn._graph.entries, "segments were not captured"
        assert not np.array_equal(seen[2][0], seen[3][0]), "the span mask did not change between graph replays"
        assert not np.array_equal(seen[2][1], seen[3][1]), "the negatives did not change between graph replays"
        assert np.array_equal(seen[3][0], seen[4][0]) and np.array_equal(seen[3][1], seen[4][1]), \
            "the same torch seed must give the same draws"
    finally:
        W.set_device_draws(False)
        graphs.set_enabled(was)


def _same_rng(before, after, n_layer_draws):
    """numpy's global generator advanced by at most the per-layer LayerDrop draws (one double = two 32-bit words each)"""
    probe = np.random.RandomState()
    probe.set_state(before)
    for _ in range(n_layer_draws + 1):
        st = probe.get_state()
        if st[2] == after[2] and np.array_equal(st[1], after[1]):
            return True
        probe.random_sample()
    return False


VQ_LOG = []  # (case, code-index flips against the oracle's free arg-max, entries)


def run_dropout_statistics(device, n_seeds=8):
    """The BENCHMARKED configuration's dropout (0.1 at the reference's five sites) has no bit-level oracle: torch's and
    this package's Philox streams differ.  Statistical parity instead: the loss averaged over `n_seeds` dropout seeds
    (same weights, input, mask, negatives, Gumbel noise) agrees with the oracle's average under torch's own dropout
    within 1 %, and both differ from the dropout-free loss in the same direction."""
    from audio8_b200 import wav2vec2 as W
    cfg = dict(d_model=256, num_heads=4, num_layers=3, d_ff=1024, final_dim=128, num_vq_vars=64, num_vq_groups=2)
    B, L, K, seed = 4, 48000, 50, 21
    sd = P.pretrain_state_dict(seed=13, **{k: v for k, v in cfg.items() if k != "num_heads"})
    model = W.create_model(dropout=0.1, dropout_input=0.1, dropout_features=0.1, **cfg)
    model.load_state_dict(sd, strict=True)
    model = model.to(device).train()
    n_vars = cfg["num_vq_vars"] * cfg["num_vq_groups"]
    loss_fn = W.create_loss(n_vars, K)
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(6)) * 0.1
    T = R.conv_out_lengths(L, R.CONV_FEATURES[16])[-1]
    np.random.seed(seed)
    tmask = R.create_mask((B, T), 0.65, 10)
    for _ in range(cfg["num_layers"]):
        np.random.random()
    Tm = int(tmask[0].sum())
    neg = R.sample_negative_indices(B, Tm, K)
    noise = gumbel_noise_like_torch(seed, (B * Tm * 2, cfg["num_vq_vars"]))
    model.quantizer.noise_override = noise.to(device)
    okw = dict(n_vars=n_vars, num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], num_groups=2, tau=0.5,
               gumbel_noise=noise)
    ours, ref = [], []
    xd = x.to(device)
    for i in range(n_seeds):
        np.random.seed(seed)
        torch.manual_seed(1000 + i)
        ours.append(loss_fn(model, xd).item())
        torch.manual_seed(2000 + i)
        with torch.no_grad():
            ref.append(R.pretrain_loss(sd, x, tmask, neg, dropout=0.1, dropout_input=0.1, dropout_features=0.1, **okw)["loss"].item())
    with torch.no_grad():
        base = R.pretrain_loss(sd, x, tmask, neg, **okw)["loss"].item()
    mo, mr = float(np.mean(ours)), float(np.mean(ref))
    assert len(set(round(v, 6) for v in ours)) > 1, "dropout did not vary with the torch seed"
    assert abs(mo - mr) <= 1e-2 * abs(mr), f"mean loss under dropout 0.1: ours {mo:.5f} vs oracle {mr:.5f} (no dropout {base:.5f})"
    return mo, mr, base


def run_acoustic_generic(device, cfg, V, B, L, S, seed=4, train=True,
                         check_grads=("proj.weight", "encoder.mask_emb", "encoder.proj_to_input.layer.weight",
                                      "encoder.encoder.transformer.encoders.0.ffn.3.layer.weight",
                                      "encoder.encoder.transformer.encoders.0.self_attn.w_Q.layer.weight"),
                         in_lens=None, tgt_lens=None, freeze_fx=True, case=None):
    """CTC fine-tuning step at any size (BASELINE configs[2] per GPU at full size): ragged utterance lengths, time and
    channel masks drawn by the product under a numpy seed and re-drawn for the oracle from the same seed (reference
    order: time mask, channel mask, one draw per layer), dropout 0, feature encoder frozen as in train.py's default."""
    from audio8_b200 import wav2vec2 as W
    from audio8_b200.ctc import ctc_loss
    CASE[0] = case or f"acoustic d={cfg.get('d_model', 768)} L={cfg.get('num_layers', 12)} B={B} x {L} V={V}"
    sd = P.acoustic_state_dict(V, seed=12, **{k: v for k, v in cfg.items() if k != "num_heads"})
    model = W.create_acoustic_model(V, dropout=0.0, freeze_fx=freeze_fx, **cfg)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model = model.to(device)
    model.freeze = False
    model.train(train)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(B, L, generator=g) * 0.1
    in_len = torch.randint(int(0.6 * L), L + 1, (B,), generator=g)
    in_len[0] = L
    if in_lens is not None:
        in_len = torch.tensor(in_lens)
    for b in range(B):
        x[b, in_len[b]:] = 0
    pad_mask = torch.arange(L)[None, :] < in_len[:, None]
    tgt_len = torch.randint(max(S // 2, 1), S + 1, (B,), generator=g)
    if tgt_lens is not None:
        tgt_len = torch.tensor(tgt_lens)
    targets = torch.full((B, S), 1, dtype=torch.long)
    for b in range(B):
        targets[b, : tgt_len[b]] = torch.randint(4, V, (int(tgt_len[b]),), generator=g)
    np.random.seed(seed)
    lp, fmask = model(x.to(device), pad_mask.to(device))
    out_len = fmask.sum(-1)
    # the loss works from the logits when it can (fused log_softmax + CTC): its gradient then lands there, not on lp
    logits_ours = getattr(lp, "a8_logits", None)
    (logits_ours if logits_ours is not None else lp).retain_grad()
    loss = ctc_loss(lp.transpose(1, 0), out_len, targets.to(device), tgt_len, blank=0, pad=1, eos=2)
    loss.backward()
    T, D = lp.shape[1], cfg.get("d_model", 768)
    tm = cm = None
    if train:
        np.random.seed(seed)
        tm = R.create_mask((B, T), 0.5, 10)
        cm = R.create_mask((B, D), 0.1, 64)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lp2, fm2, logits2 = R.acoustic_forward(sdg, x, pad_mask, cfg.get("num_heads", 12), cfg.get("num_layers", 12), tm, cm,
                                           freeze_fx=freeze_fx, return_logits=True)
    assert (fm2.sum(-1).numpy() == out_len.cpu().numpy()).all(), "frame lengths differ"
    # CTC occupancies are exponentially sensitive to the sequence of log-probs (an untrained model spreads its mass
    # over ~10^100 alignments: the bf16-sized log-prob differences, 0.3 % rms, move dL/dlogprob by ~25 % at T=749 even
    # though the loss agrees to 0.07 %), so at this size parity is checked piecewise, every piece on identical inputs:
    #   (1) log-probs and loss against the oracle;
    #   (2) the CTC kernel against float64 CTC on the ORACLE's log-probs (ATen's fp32 CTC is the less accurate one);
    #   (3) the network's backward against the oracle's backward driven by the SAME dL/dlogits (ours).
    valid = fm2[..., None].expand_as(lp2)
    act_close(torch.where(valid, lp.detach().float().cpu(), torch.zeros(())), torch.where(valid, lp2.detach(), torch.zeros(())),
              "log-probs (valid frames)", tol=5e-2)
    lp64 = lp2.detach().double().requires_grad_(True)
    loss2 = ref_ctc.ctc_loss_reference(lp64.transpose(1, 0), out_len.cpu(), targets, tgt_len, 0, 1, 2)
    loss2.backward()
    assert abs(loss.item() - loss2.item()) <= 2e-2 * abs(loss2.item()), (loss.item(), loss2.item())
    lp_same = lp2.detach().to(device).requires_grad_(True)
    loss_same = ctc_loss(lp_same.transpose(1, 0), out_len, targets.to(device), tgt_len, blank=0, pad=1, eos=2)
    loss_same.backward()
    assert abs(loss_same.item() - loss2.item()) <= 2e-5 * abs(loss2.item()), (loss_same.item(), loss2.item())
    grad_close(lp_same.grad.cpu(), lp64.grad, "CTC gradient on identical log-probs", cos_min=0.99999, rel_max=2e-3)
    if logits_ours is not None:
        logits2.backward(logits_ours.grad.detach().float().cpu())
    else:  # e.g. a head width that is not a multiple of 8: the sliced logits are not the contiguous tensor the loss wants
        lp2.backward(lp.grad.detach().float().cpu())
    got = dict(model.named_parameters())
    bad = []
    for k in check_grads:
        try:
            grad_close(got[k].grad, sdg[k].grad, "grad " + k, **grad_tol(k))
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, "; ".join(bad)
    for k, p_ in got.items():
        if "feature_extractor" in k and freeze_fx:
            assert p_.grad is None, f"frozen feature encoder got a gradient: {k}"
    return loss.item(), loss2.item()
