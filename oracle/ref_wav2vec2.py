"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch, functional, fp32/fp64) of the reference's
wav2vec2 training hot path.  Never imported by the product package `audio8_b200/`.

It works on a reference-format `state_dict` (the on-disk contract) instead of `nn.Module`s, so
it checks key names/shapes too, and it exposes every random draw (time mask, Gumbel noise, negative
indices) as an explicit argument so parity tests can share them with the CUDA path.

Pinned by `oracle/gen_golden.py` against the UNMODIFIED `/root/reference/audio8/wav2vec2.py` run in the dev
container (fixtures in `tests/golden/`).  The `eight_mile` pieces (transformer layer order, LN eps 1e-6, pad
split 63/64) are recalled — parity unpinned at that boundary (see `eight_mile_compat.py`).
Dropout (torch's own F.dropout, at the reference's five sites) and LayerDrop (an explicit list of active layers)
are optional arguments, off by default: exact parity runs use p = 0 (SURVEY §8c protocol); the statistical
dropout test and the GPU-incumbent timing in bench.py switch them on.  Every function follows the device of its
tensor arguments (the GPU-incumbent leg of bench.py runs this same code on `cuda`).
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

CONV_FEATURES = {  # reference wav2vec2.py:26-29
    16: [(512, 10, 5), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 2, 2), (512, 2, 2)],
    8: [(512, 10, 5), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 2, 2), (512, 2, 2)],
}
XE_WGT = 0.1  # wav2vec2.py:34
DIVERSITY_WGT = 10  # wav2vec2.py:35
LN_EPS_TORCH = 1e-5  # torch.nn.LayerNorm default (wav2vec2.py:623,904)
LN_EPS_8MILE = 1e-6  # eight_mile TransformerEncoder layer_norm_eps [RECALLED]


# ------------------------------------------------------------------------------------------------
# host-side integer work (bit-exact contracts)
# ------------------------------------------------------------------------------------------------
def create_mask(shape, p_start=0.65, mask_length=10, rng=np.random):
    """Span mask, every row subsampled to the batch-min count.  Follows wav2vec2.py:189-216, including the
    order of numpy global-RNG draws: rand() once, choice() per row, choice() per over-long row."""
    bsz, T = shape
    mask = np.zeros((bsz, T), dtype=bool)
    num_mask = int(p_start * T / float(mask_length) + rng.rand())
    if num_mask == 0:
        return mask
    rows = []
    for _ in range(bsz):
        span = mask_length
        if T - span <= num_mask:  # wav2vec2.py:200-201
            span = T - num_mask - 1
        starts = rng.choice(T - span, num_mask, replace=False)
        idx = (starts[:, None] + np.arange(mask_length)[None, :]).reshape(-1)
        rows.append(np.unique(idx[idx < T]))
    keep = min(len(r) for r in rows)
    for i, r in enumerate(rows):
        if len(r) > keep:
            r = rng.choice(r, keep, replace=False)
        mask[i, r] = True
    return mask


def sample_negative_indices(B, Tm, n_negatives, rng=np.random):
    """wav2vec2.py:959-973: K distractors per masked step drawn from the other masked steps of the same
    utterance; returned already offset by b*Tm, shape [B, K*Tm] int64 (row-major over (t, k))."""
    own_t = np.repeat(np.arange(Tm), n_negatives)[None, :]
    idx = rng.randint(0, Tm - 1, (B, n_negatives * Tm))
    idx = np.where(idx >= own_t, idx + 1, idx)
    return idx + (np.arange(B) * Tm)[:, None]


def frame_mask_from_sample_mask(pad_mask, T):
    """wav2vec2.py:703-708: trim L % T samples, chunk into T groups, a frame is valid iff all samples are."""
    B, L = pad_mask.shape
    extra = L % T
    if extra > 0:
        pad_mask = pad_mask[:, :-extra]
    return pad_mask.reshape(B, T, -1).all(-1)


def conv_out_lengths(L, conv_features):
    out = []
    for (_, k, s) in conv_features:
        L = (L - k) // s + 1
        out.append(L)
    return out


# ------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------
class _RoundBF16(torch.autograd.Function):
    """value and gradient rounded to bf16 (computed in fp32 on both sides): one STORED activation of the implementation
    under test"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


STORAGE_EMULATION = [False]


def set_storage_emulation(flag):
    """Test aid (not in the reference): round every activation the CUDA path STORES in bf16 — conv / linear / LayerNorm
    outputs, residual sums, attention probabilities and context — and the gradients at the same points, while all
    arithmetic stays fp32.  The distance between this run and the plain fp32 oracle is the noise floor of the storage
    format itself, which a correct kernel cannot beat and a wrong one exceeds."""
    STORAGE_EMULATION[0] = bool(flag)


def _q(x):
    return _RoundBF16.apply(x) if STORAGE_EMULATION[0] else x


def _lin(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def _ln(sd, key, x, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps)


def conv_feature_extractor(sd, x, prefix="feature_extractor.", conv_features=CONV_FEATURES[16]):
    """wav2vec2.py:399-456: [B,L] -> [B,512,T]; layer 0 has GroupNorm(512,512) (per (b,channel) over time),
    every layer exact-erf GELU, no conv bias, Dropout(0)."""
    h = x.unsqueeze(1)
    for i, (_, k, s) in enumerate(conv_features):
        h = F.conv1d(h, sd[f"{prefix}conv_layers.{i}.0.weight"], None, stride=s)
        if i == 0:
            C = h.shape[1]
            h = F.group_norm(h, C, sd[f"{prefix}conv_layers.0.2.weight"], sd[f"{prefix}conv_layers.0.2.bias"], 1e-5)
        h = _q(F.gelu(h))
    return h


def pos_conv_weight(sd, prefix):
    """weight_norm(dim=2) (wav2vec2.py:609): w = g * v / ||v|| with the norm over (out, in/groups) per tap."""
    v = sd[prefix + "pos_conv.conv.1.weight_v"]
    g = sd[prefix + "pos_conv.conv.1.weight_g"]
    return g * v / v.norm(2, dim=(0, 1), keepdim=True)


def _drop(x, p):
    return F.dropout(x, p, True) if p > 0 else x


def transformer_layer(sd, pre, x, num_heads, key_mask=None, pdrop=0.0):
    """eight_mile TransformerEncoder with layer_norms_after=True [RECALLED; structure per wav2vec2.py:110-126]:
    x = ln2(x + MHA(x)); x = ln1(x + FFN(x)).  key_mask: bool [B,T], False = padded key."""
    B, T, D = x.shape
    dk = D // num_heads
    q = _q(_lin(sd, pre + "self_attn.w_Q.layer", x)).view(B, T, num_heads, dk).transpose(1, 2)
    k = _q(_lin(sd, pre + "self_attn.w_K.layer", x)).view(B, T, num_heads, dk).transpose(1, 2)
    v = _q(_lin(sd, pre + "self_attn.w_V.layer", x)).view(B, T, num_heads, dk).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(dk)
    if key_mask is not None:
        s = s.masked_fill(~key_mask[:, None, None, :], -1e9)
    a = _q(_q(_drop(torch.softmax(s, -1), pdrop)) @ v)  # eight_mile SeqScaledDotProductAttention: dropout on the probabilities
    a = a.transpose(1, 2).reshape(B, T, D)
    x = _q(_ln(sd, pre + "ln2", _q(x + _drop(_q(_lin(sd, pre + "self_attn.w_O.layer", a)), pdrop)), LN_EPS_8MILE))
    f = _q(_lin(sd, pre + "ffn.3.layer", _q(F.gelu(_lin(sd, pre + "ffn.0.layer", x)))))
    return _q(_ln(sd, pre + "ln1", _q(x + _drop(f, pdrop)), LN_EPS_8MILE))


def audio_transformer_encoder(sd, x, num_heads, num_layers, prefix="encoder.", pad_mask=None, groups=16, pdrop=0.0,
                              active_layers=None):
    """wav2vec2.py:629-646: zero padded frames, x += gelu(pos_conv(x)), LN, dropout, transformer stack.
    active_layers: LayerDrop outcome, one bool per layer (eight_mile skips a layer when its draw < layer_drop)."""
    if pad_mask is not None:
        x = x.masked_fill(~pad_mask[..., None], 0.0)
    w = pos_conv_weight(sd, prefix)
    k = w.shape[-1]
    end_pad = k // 2
    start_pad = end_pad - 1 if k % 2 == 0 else end_pad
    xc = F.conv1d(F.pad(x.transpose(1, 2), (start_pad, end_pad)), w, sd[prefix + "pos_conv.conv.1.bias"], groups=groups)
    x = _q(x + F.gelu(xc).transpose(1, 2))
    x = _q(_drop(_ln(sd, prefix + "ln", x, LN_EPS_TORCH), pdrop))
    for i in range(num_layers):
        if active_layers is not None and not active_layers[i]:
            continue
        x = transformer_layer(sd, f"{prefix}transformer.encoders.{i}.", x, num_heads, pad_mask, pdrop)
    return x


def gumbel_quantizer(sd, y, num_groups, tau=0.5, gumbel_noise=None, prefix="quantizer.", force_idx=None, force_z=None):
    """wav2vec2.py:547-576.  y [B,Tm,512] -> (q [B,Tm,G*var_dim], prob_ppl, argmax indices [B*Tm*G]).
    training: gumbel_noise [B*Tm*G, V] (= -log(Exp(1)), as F.gumbel_softmax draws it); eval: None.
    force_idx (test aid, not in the reference): use these code indices instead of the arg-max, so that a
    bf16 near-tie flip upstream does not pollute every downstream comparison; flips are counted separately."""
    B, Tm, _ = y.shape
    z = _lin(sd, prefix + "weight_proj", y).reshape(B * Tm * num_groups, -1).float()
    if force_z is not None:
        # test aid, not in the reference: evaluate at the logit VALUES of the implementation under test (the autograd
        # graph stays the reference's).  |z| ~ 20 (weight_proj ~ N(0,1), wav2vec2.py:486) and tau = 0.5 make softmax(u) move
        # by ~10 % for a 2^-9 relative change of the features, so every gradient upstream of the quantizer inherits that
        # sensitivity; pinning z separates it from the arithmetic of the backward kernels
        z = z + (torch.as_tensor(force_z).reshape(z.shape).to(z) - z).detach()
    V = z.shape[-1]
    avg_probs = torch.softmax(z, -1).mean(0)  # pooled over groups: shape [V]  (wav2vec2.py:554)
    if gumbel_noise is not None:
        u = (z + gumbel_noise) / tau
        soft = torch.softmax(u, -1)
        k = soft.argmax(-1) if force_idx is None else torch.as_tensor(force_idx).long().to(z.device)
        hard = torch.zeros_like(z).scatter_(-1, k[:, None], 1.0)
        onehot = hard - soft.detach() + soft  # straight-through (torch F.gumbel_softmax hard=True)
    else:
        k = z.argmax(-1) if force_idx is None else torch.as_tensor(force_idx).long().to(z.device)
        onehot = torch.zeros_like(z).scatter_(-1, k[:, None], 1.0)
    ppl = torch.exp(-torch.sum(avg_probs * torch.log(avg_probs + 1e-7)))  # wav2vec2.py:565
    vars_ = sd[prefix + "vars"]  # [1, G*V, var_dim]
    q = (onehot.reshape(B * Tm, num_groups * V, 1) * vars_).reshape(B * Tm, num_groups, V, -1).sum(-2)
    return q.reshape(B, Tm, -1), ppl, k


def contrastive_loss(x_masked, y, neg_idx, ppl, n_vars):
    """wav2vec2.py:377-392.  x_masked [B,Tm,C] (context outputs at masked steps), y [B,Tm,C] (quantized),
    neg_idx [B,K*Tm] (already offset).  loss = 0.1*CE(cos-sim logits, class 0) + 10*(n_vars-ppl)/n_vars."""
    B, Tm, C = y.shape
    K = neg_idx.shape[1] // Tm
    negs = y.reshape(-1, C)[torch.as_tensor(neg_idx).reshape(-1).to(y.device)].view(B, Tm, K, C).permute(2, 0, 1, 3)
    targets = torch.cat([y.unsqueeze(0), negs], 0)
    logits = torch.cosine_similarity(x_masked.unsqueeze(0), targets, dim=-1)  # [K+1,B,Tm]
    logits = logits.transpose(2, 0).reshape(-1, K + 1)
    ce = F.cross_entropy(logits, torch.zeros(logits.shape[0], dtype=torch.long, device=logits.device))
    return XE_WGT * ce + DIVERSITY_WGT * (n_vars - ppl) / n_vars, ce


# ------------------------------------------------------------------------------------------------
# whole-model forwards
# ------------------------------------------------------------------------------------------------
def pretrain_forward(sd, x, time_mask, num_heads=12, num_layers=12, num_groups=2, tau=0.5, gumbel_noise=None,
                     conv_features=CONV_FEATURES[16], force_idx=None, dropout=0.0, dropout_input=0.0,
                     dropout_features=0.0, active_layers=None, force_z=None):
    """Wav2Vec2Model.forward (wav2vec2.py:927-952) with the time mask supplied.  Returns a dict of stages."""
    fx = conv_feature_extractor(sd, x, conv_features=conv_features).transpose(1, 2)
    feats = _ln(sd, "layer_norm", fx, LN_EPS_TORCH)
    unmasked = feats  # the quantizer branch reads the fp32 copy of the LayerNorm output in the CUDA path as well
    h = _q(_drop(_lin(sd, "proj_to_input.layer", _q(feats)), dropout_input))
    unmasked = _drop(unmasked, dropout_features)
    B, T, _ = h.shape
    tm = torch.as_tensor(time_mask).to(h.device)
    h = torch.where(tm[..., None], sd["mask_emb"].expand_as(h), h)
    y_in = unmasked[tm].view(B, -1, unmasked.shape[-1])
    enc = audio_transformer_encoder(sd, h, num_heads, num_layers, pdrop=dropout, active_layers=active_layers)
    q, ppl, k = gumbel_quantizer(sd, y_in, num_groups, tau, gumbel_noise, force_idx=force_idx, force_z=force_z)
    y = _lin(sd, "project_q.layer", q)
    xo = _lin(sd, "final_proj.layer", enc)
    return dict(fx=fx, features=feats, y_in=y_in, enc=enc, q=q, ppl=ppl, vq_idx=k, y=y, x=xo)


def pretrain_loss(sd, x, time_mask, neg_idx, n_vars=640, **kw):
    """Wav2Vec2Loss.__call__ (wav2vec2.py:377-392) with the random draws supplied."""
    st = pretrain_forward(sd, x, time_mask, **kw)
    tm = torch.as_tensor(time_mask).to(x.device)
    B = x.shape[0]
    xm = st["x"][tm].view(B, -1, st["x"].shape[-1])
    loss, ce = contrastive_loss(xm, st["y"], neg_idx, st["ppl"], n_vars)
    st.update(loss=loss, ce=ce, x_masked=xm)
    return st


def acoustic_forward(sd, x, pad_mask, num_heads=12, num_layers=12, time_mask=None, channel_mask=None,
                     conv_features=CONV_FEATURES[16], dropout=0.0, dropout_input=0.0, active_layers=None,
                     freeze_fx=False, return_logits=False):
    """Wav2Vec2AcousticModel.forward (wav2vec2.py:765-770) over Wav2Vec2Encoder.forward (:696-723).
    time_mask / channel_mask: the training-time masks (None = eval).  Returns (log_probs [B,T,V], frame_mask)."""
    p = "encoder."
    with torch.no_grad() if freeze_fx else torch.enable_grad():
        fx = conv_feature_extractor(sd, x, prefix=p + "feature_extractor.", conv_features=conv_features).transpose(1, 2)
    feats = _ln(sd, p + "layer_norm", fx, LN_EPS_TORCH)
    T = feats.shape[1]
    fmask = frame_mask_from_sample_mask(pad_mask, T) if pad_mask is not None else None
    h = _drop(_lin(sd, p + "proj_to_input.layer", feats), dropout_input)
    if time_mask is not None:
        h = torch.where(torch.as_tensor(time_mask).to(h.device)[..., None], sd[p + "mask_emb"].expand_as(h), h)
    if channel_mask is not None:
        h = h.masked_fill(torch.as_tensor(channel_mask).to(h.device)[:, None, :], 0.0)
    enc = audio_transformer_encoder(sd, h, num_heads, num_layers, prefix=p + "encoder.", pad_mask=fmask, pdrop=dropout,
                                    active_layers=active_layers)
    logits = _lin(sd, "proj", enc)
    if return_logits:
        return F.log_softmax(logits, -1), fmask, logits
    return F.log_softmax(logits, -1), fmask
