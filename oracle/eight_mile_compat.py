"""TEST INFRASTRUCTURE ONLY — restatement of the `eight_mile` (mead-baseline) layers the reference calls.

The reference (`/root/reference/audio8/wav2vec2.py:7-21`, `ctc.py:5`, `train.py:15-17`) imports these from the
third-party package `mead-baseline`, which is unpinned (`/root/reference/setup.cfg:41`), not vendored and not
installable here (no network).  This file restates their published behaviour in plain PyTorch so that
(a) the unmodified reference files import through `oracle/shim/`, and (b) `oracle/ref_wav2vec2.py` has the same
building blocks.  The module tree (attribute names) is pinned by the reference's own fairseq key maps
(`wav2vec2.py:39-151`); numerical details (LN eps 1e-6, pad split 63/64, one numpy draw per layer) are recalled —
"parity unpinned" at this boundary.  Never imported by the product package.
"""
import math
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class Offsets:
    """eight_mile.utils.Offsets; `train.py:22-27` re-points GO/PAD at import time."""

    PAD, GO, EOS, UNK, OFFSET = 0, 1, 2, 3, 4
    VALUES = ["<PAD>", "<GO>", "<EOS>", "<UNK>"]


def get_activation(name):
    if name is None or name == "ident":
        return nn.Identity()
    if name == "gelu":
        return nn.GELU()
    if name == "relu":
        return nn.ReLU()
    if name == "tanh":
        return nn.Tanh()
    raise ValueError(name)


def pytorch_linear(in_sz, out_sz, unif=0, initializer=None, bias=True):
    l = nn.Linear(in_sz, out_sz, bias=bias)
    if unif > 0:
        l.weight.data.uniform_(-unif, unif)
    elif initializer == "ortho":
        nn.init.orthogonal_(l.weight)
    elif initializer == "he" or initializer == "kaiming":
        nn.init.kaiming_uniform_(l.weight)
    else:
        nn.init.xavier_uniform_(l.weight)
    if bias:
        l.bias.data.zero_()
    return l


def pytorch_conv1d(in_channels, out_channels, fsz, unif=0, padding=0, initializer=None, stride=1, bias=True, groups=1):
    c = nn.Conv1d(in_channels, out_channels, fsz, padding=padding, stride=stride, bias=bias, groups=groups)
    if unif > 0:
        c.weight.data.uniform_(-unif, unif)
    elif initializer == "ortho":
        nn.init.orthogonal_(c.weight)
    elif initializer == "he" or initializer == "kaiming":
        nn.init.kaiming_uniform_(c.weight)
    elif initializer == "normal":
        nn.init.normal_(c.weight, mean=0, std=unif)
    else:
        nn.init.xavier_uniform_(c.weight)
    if bias:
        c.bias.data.zero_()
    return c


class Dense(nn.Module):
    def __init__(self, insz, outsz, activation=None, unif=0, initializer=None):
        super().__init__()
        self.layer = pytorch_linear(insz, outsz, unif, initializer)
        self.activation = get_activation(activation)
        self.output_dim = outsz

    def forward(self, x):
        return self.activation(self.layer(x))


class Conv1DSame(nn.Module):
    """'same' conv for even or odd kernels: pad (k//2 - 1, k//2) when k is even."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=True, groups=1, unif=0.0, initializer=None, activation=None):
        super().__init__()
        end_pad = kernel_size // 2
        start_pad = end_pad - 1 if kernel_size % 2 == 0 else end_pad
        self.conv = nn.Sequential(
            nn.ConstantPad1d((start_pad, end_pad), 0.0),
            pytorch_conv1d(in_channels, out_channels, kernel_size, unif=unif, initializer=initializer, bias=bias, groups=groups),
            get_activation(activation),
        )

    def forward(self, x):
        return self.conv(x)


class SeqScaledDotProductAttention(nn.Module):
    def __init__(self, pdrop=0.1):
        super().__init__()
        self.dropout = nn.Dropout(pdrop)

    def forward(self, query, key, value, mask=None):
        d_k = query.size(-1)
        scores = torch.matmul(query, key.transpose(-2, -1)) / math.sqrt(d_k)
        if mask is not None:
            scores = scores.masked_fill(mask == False, -1e9)  # noqa: E712
        a = F.softmax(scores, dim=-1)
        return torch.matmul(self.dropout(a), value)


class MultiHeadedAttention(nn.Module):
    def __init__(self, num_heads, d_model, dropout=0.1, scale=True, d_k=None):
        super().__init__()
        self.d_k = d_model // num_heads if d_k is None else d_k
        self.h = num_heads
        self.w_Q = Dense(d_model, self.d_k * self.h)
        self.w_K = Dense(d_model, self.d_k * self.h)
        self.w_V = Dense(d_model, self.d_k * self.h)
        self.w_O = Dense(self.d_k * self.h, d_model)
        self.attn_fn = SeqScaledDotProductAttention(dropout)

    def forward(self, qkvm):
        query, key, value, mask = qkvm
        B = query.size(0)
        q = self.w_Q(query).view(B, -1, self.h, self.d_k).transpose(1, 2)
        k = self.w_K(key).view(B, -1, self.h, self.d_k).transpose(1, 2)
        v = self.w_V(value).view(B, -1, self.h, self.d_k).transpose(1, 2)
        x = self.attn_fn(q, k, v, mask=mask)
        x = x.transpose(1, 2).contiguous().view(B, -1, self.h * self.d_k)
        return self.w_O(x)


class FFN(nn.Sequential):
    """Sequential(Dense, act, Dropout, Dense): keys ffn.0.layer / ffn.3.layer (`wav2vec2.py:121-124`)."""

    def __init__(self, d_model, activation="gelu", d_ff=None, pdrop=0.0):
        d_ff = d_ff if d_ff else 4 * d_model
        super().__init__(Dense(d_model, d_ff), get_activation(activation), nn.Dropout(pdrop), Dense(d_ff, d_model))


class TransformerEncoder(nn.Module):
    def __init__(self, num_heads, d_model, pdrop, scale=True, activation_type="gelu", d_ff=None, d_k=None,
                 ffn_pdrop=0.0, layer_norms_after=False, layer_norm_eps=1e-6):
        super().__init__()
        self.layer_norms_after = layer_norms_after
        self.d_model = d_model
        self.d_ff = d_ff if d_ff is not None else 4 * d_model
        self.self_attn = MultiHeadedAttention(num_heads, d_model, pdrop, scale=scale, d_k=d_k)
        self.ffn = FFN(d_model, activation_type, self.d_ff, ffn_pdrop)
        self.ln1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.ln2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.dropout = nn.Dropout(pdrop)

    def forward(self, inputs):
        x, mask = inputs
        if not self.layer_norms_after:
            x = self.ln1(x)
        h = self.self_attn((x, x, x, mask))
        x = x + self.dropout(h)
        x = self.ln2(x)
        x = x + self.dropout(self.ffn(x))
        if self.layer_norms_after:
            x = self.ln1(x)
        return x


class TransformerEncoderStack(nn.Module):
    def __init__(self, num_heads, d_model, pdrop, scale=True, layers=1, activation="gelu", d_ff=None, d_k=None,
                 rpr_k=None, ffn_pdrop=0.0, layer_norms_after=False, layer_norm_eps=1e-6, layer_drop=0.0, **kwargs):
        super().__init__()
        self.encoders = nn.ModuleList()
        self.ln = nn.Identity() if layer_norms_after else nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.output_dim = d_model
        self.layer_drop = layer_drop
        for _ in range(layers):
            self.encoders.append(
                TransformerEncoder(num_heads, d_model, pdrop, scale, activation, d_ff, d_k, ffn_pdrop=ffn_pdrop,
                                   layer_norms_after=layer_norms_after, layer_norm_eps=layer_norm_eps))

    def forward(self, inputs):
        x, mask = inputs
        for layer in self.encoders:
            pdrop = np.random.random()  # one numpy draw per layer even when layer_drop == 0
            if not self.training or (pdrop >= self.layer_drop):
                x = layer((x, mask))
        return self.ln(x)


def sequence_mask(lengths, max_len=-1):
    lens = lengths.cpu()
    if max_len < 0:
        max_len = int(torch.max(lens).item())
    row = torch.arange(0, max_len).type_as(lens).view(1, -1)
    col = lens.view(-1, 1)
    return (row < col).to(lengths.device)


def sequence_mask_mxlen(lengths, max_len):
    return sequence_mask(lengths, max_len)


class _Inert(nn.Module):
    """Stand-in for eight_mile symbols the reference imports but the hot path never constructs."""

    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("inert eight_mile stand-in (outside the hot path)")


PassThru = MaxPool1D = MeanPool1D = TwoHeadConcat = SingleHeadReduction = BasicDualEncoderModel = _Inert
EmbeddingsStack = TransformerDecoderStack = WeightTieDense = _Inert


def subsequent_mask(size):
    return torch.tril(torch.ones(1, 1, size, size)).bool()


def load_tlm_npz(*a, **k):
    raise NotImplementedError
