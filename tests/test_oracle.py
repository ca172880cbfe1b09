"""The CPU oracle against the fixtures generated from the UNMODIFIED reference (oracle/gen_golden.py)."""
import os

import numpy as np
import torch

import ref_params as P
import ref_wav2vec2 as R
from model_cases import GOLD_DIR, TINY_PRE, gumbel_noise_like_torch

HOST = np.load(os.path.join(GOLD_DIR, "host_integer.npz"))


def test_create_mask_bit_exact():
    from audio8_b200.wav2vec2 import create_mask
    for name in ("c2", "c1", "chan", "short"):
        B, T, p1000, ln, seed = (int(v) for v in HOST[f"mask_{name}_cfg"])
        want = np.unpackbits(HOST[f"mask_{name}"], axis=1)[:, :T].astype(bool)
        for fn in (R.create_mask, create_mask):  # oracle and product share the reference's draw order
            np.random.seed(seed)
            assert (fn((B, T), p1000 / 1000.0, ln) == want).all(), (name, fn.__module__)
        assert len(set(want.sum(1))) == 1  # every row has the same number of masked steps


def test_negative_indices_bit_exact():
    from audio8_b200.wav2vec2 import Sampler
    for name in ("c2", "small"):
        B, Tm, K, seed = (int(v) for v in HOST[f"neg_{name}_cfg"])
        for fn in (lambda: R.sample_negative_indices(B, Tm, K), lambda: Sampler(K).indices(B, Tm)):
            np.random.seed(seed)
            idx = fn()
            chk = np.array([idx.sum(), (idx * np.arange(idx.size).reshape(idx.shape) % 65521).sum()])
            assert (chk == HOST[f"neg_{name}_sum"]).all(), name
            own = np.repeat(np.arange(Tm), K)[None, :] + (np.arange(B) * Tm)[:, None]
            assert (idx != own).all(), "a negative may never be the positive (wav2vec2.py:969)"
            assert ((idx // Tm) == np.arange(B)[:, None]).all(), "negatives stay inside the utterance"
    np.random.seed(int(HOST["neg_small_cfg"][3]))
    assert (R.sample_negative_indices(*(int(v) for v in HOST["neg_small_cfg"][:3])) == HOST["neg_small"]).all()


def test_pretrain_oracle_matches_reference_fixture():
    gold = np.load(os.path.join(GOLD_DIR, "pretrain_tiny.npz"))
    B, L, K, seed, wseed, xseed = (int(v) for v in gold["cfg"])
    cfg = dict(TINY_PRE)
    sd = P.pretrain_state_dict(seed=wseed, **{k: v for k, v in cfg.items() if k != "num_heads"})
    x = torch.randn(B, L, generator=torch.Generator().manual_seed(xseed)) * 0.1
    for mode in ("train", "eval"):
        T = R.conv_out_lengths(L, R.CONV_FEATURES[16])[-1]
        tmask = np.unpackbits(gold[mode + "_time_mask"], axis=1)[:, :T].astype(bool)
        Tm = int(tmask[0].sum())
        noise = gumbel_noise_like_torch(seed, (B * Tm * 2, cfg["num_vq_vars"])) if mode == "train" else None
        with torch.no_grad():
            st = R.pretrain_loss(sd, x, tmask, gold[mode + "_neg_idx"].astype(np.int64), n_vars=48, num_heads=2,
                                 num_layers=2, num_groups=2, tau=0.5, gumbel_noise=noise)
        want = gold[mode + "_loss"]
        assert abs(st["loss"].item() - want[0]) < 2e-5 and abs(st["ce"].item() - want[1]) < 2e-5
        assert abs(st["ppl"].item() - want[2]) < 1e-4
        assert (st["vq_idx"].numpy() == gold[mode + "_vq_idx"]).all()
        np.testing.assert_allclose(st["x"].numpy()[:, ::7, ::5], gold[mode + "_x"], atol=2e-5)
        np.testing.assert_allclose(st["fx"].numpy()[:, ::6, ::37], gold[mode + "_fx"], atol=2e-5)


def test_state_dict_contract():
    """product modules expose exactly the reference's keys and shapes (fixtures' weights load with strict=True)"""
    from audio8_b200 import wav2vec2 as W
    sd = P.pretrain_state_dict(seed=0, d_model=128, num_layers=1, d_ff=256, final_dim=64, num_vq_vars=24)
    m = W.create_model(d_model=128, num_heads=2, num_layers=1, d_ff=256, final_dim=64, num_vq_vars=24, foo="swallowed")
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    sa = P.acoustic_state_dict(32, seed=0, d_model=128, num_layers=1, d_ff=256)
    a = W.create_acoustic_model(32, d_model=128, num_heads=2, num_layers=1, d_ff=256, bar=1)
    assert {k: tuple(v.shape) for k, v in a.state_dict().items()} == {k: tuple(v.shape) for k, v in sa.items()}
    full = W.create_model()
    assert sum(p.numel() for p in full.parameters()) == 95044608  # published wav2vec2-base size
