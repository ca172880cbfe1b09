"""TEST INFRASTRUCTURE ONLY — deterministic reference-format `state_dict`s for parity tests.

The golden fixtures under `tests/golden/` cannot carry weights (a 512-channel conv stack is 16 MB), so both the
fixture generator (`gen_golden.py`, which loads these into the UNMODIFIED reference with `strict=True`) and the
tests rebuild the weights from a seed with the functions below.  Key names and shapes are the reference's
on-disk contract (`/root/reference/audio8/wav2vec2.py:39-151`).
"""
import math
import torch

from ref_wav2vec2 import CONV_FEATURES


def _rand(g, shape, scale):
    return (torch.rand(shape, generator=g) * 2 - 1) * scale


def _randn(g, shape, std):
    return torch.randn(shape, generator=g) * std


def _encoder_side(sd, g, p, d_model, num_layers, d_ff, conv_features, pos_k=128, groups=16):
    """keys shared by Wav2Vec2Model (p='') and Wav2Vec2Encoder (p='encoder.')"""
    cin = 1
    for i, (c, k, _) in enumerate(conv_features):
        sd[f"{p}feature_extractor.conv_layers.{i}.0.weight"] = _rand(g, (c, cin, k), math.sqrt(3.0 / (cin * k)))
        cin = c
    sd[f"{p}feature_extractor.conv_layers.0.2.weight"] = 1.0 + _randn(g, (conv_features[0][0],), 0.1)
    sd[f"{p}feature_extractor.conv_layers.0.2.bias"] = _randn(g, (conv_features[0][0],), 0.1)
    fx = conv_features[-1][0]
    sd[f"{p}layer_norm.weight"] = 1.0 + _randn(g, (fx,), 0.1)
    sd[f"{p}layer_norm.bias"] = _randn(g, (fx,), 0.1)
    sd[f"{p}proj_to_input.layer.weight"] = _rand(g, (d_model, fx), math.sqrt(6.0 / (d_model + fx)))
    sd[f"{p}proj_to_input.layer.bias"] = _randn(g, (d_model,), 0.02)
    sd[f"{p}mask_emb"] = torch.rand((d_model,), generator=g)
    e = f"{p}encoder."
    sd[e + "pos_conv.conv.1.bias"] = _randn(g, (d_model,), 0.02)
    sd[e + "pos_conv.conv.1.weight_g"] = 0.5 + torch.rand((1, 1, pos_k), generator=g)
    sd[e + "pos_conv.conv.1.weight_v"] = _randn(g, (d_model, d_model // groups, pos_k), math.sqrt(4.0 / (pos_k * d_model)))
    sd[e + "ln.weight"] = 1.0 + _randn(g, (d_model,), 0.1)
    sd[e + "ln.bias"] = _randn(g, (d_model,), 0.1)
    for i in range(num_layers):
        t = f"{e}transformer.encoders.{i}."
        for w in ("w_Q", "w_K", "w_V", "w_O"):
            sd[f"{t}self_attn.{w}.layer.weight"] = _rand(g, (d_model, d_model), math.sqrt(3.0 / d_model))
            sd[f"{t}self_attn.{w}.layer.bias"] = _randn(g, (d_model,), 0.02)
        sd[t + "ffn.0.layer.weight"] = _rand(g, (d_ff, d_model), math.sqrt(6.0 / (d_model + d_ff)))
        sd[t + "ffn.0.layer.bias"] = _randn(g, (d_ff,), 0.02)
        sd[t + "ffn.3.layer.weight"] = _rand(g, (d_model, d_ff), math.sqrt(6.0 / (d_model + d_ff)))
        sd[t + "ffn.3.layer.bias"] = _randn(g, (d_model,), 0.02)
        for ln in ("ln1", "ln2"):
            sd[t + ln + ".weight"] = 1.0 + _randn(g, (d_model,), 0.1)
            sd[t + ln + ".bias"] = _randn(g, (d_model,), 0.1)


def pretrain_state_dict(seed=0, d_model=768, num_layers=12, d_ff=None, final_dim=256, num_vq_vars=320,
                        num_vq_groups=2, sample_rate=16):
    """Keys of `Wav2Vec2Model` (wav2vec2.py:871-925)."""
    g = torch.Generator().manual_seed(seed)
    d_ff = d_ff or 4 * d_model
    cf = CONV_FEATURES[sample_rate]
    sd = {}
    _encoder_side(sd, g, "", d_model, num_layers, d_ff, cf)
    fx = cf[-1][0]
    sd["quantizer.vars"] = torch.rand((1, num_vq_groups * num_vq_vars, final_dim // num_vq_groups), generator=g)
    sd["quantizer.weight_proj.weight"] = _randn(g, (num_vq_groups * num_vq_vars, fx), 1.0)
    sd["quantizer.weight_proj.bias"] = _randn(g, (num_vq_groups * num_vq_vars,), 0.02)
    sd["project_q.layer.weight"] = _rand(g, (final_dim, final_dim), math.sqrt(3.0 / final_dim))
    sd["project_q.layer.bias"] = _randn(g, (final_dim,), 0.02)
    sd["final_proj.layer.weight"] = _rand(g, (final_dim, d_model), math.sqrt(6.0 / (d_model + final_dim)))
    sd["final_proj.layer.bias"] = _randn(g, (final_dim,), 0.02)
    return sd


def acoustic_state_dict(num_labels, seed=0, d_model=768, num_layers=12, d_ff=None, sample_rate=16):
    """Keys of `Wav2Vec2AcousticModel` (wav2vec2.py:726-770)."""
    g = torch.Generator().manual_seed(seed)
    d_ff = d_ff or 4 * d_model
    sd = {}
    _encoder_side(sd, g, "encoder.", d_model, num_layers, d_ff, CONV_FEATURES[sample_rate])
    sd["proj.weight"] = _rand(g, (num_labels, d_model), math.sqrt(6.0 / (d_model + num_labels)))
    sd["proj.bias"] = _randn(g, (num_labels,), 0.02)
    return sd
