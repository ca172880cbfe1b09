"""Input feed + fairseq checkpoint import against golden vectors written from the unmodified reference
(`oracle/gen_golden_feed.py`): batches bit-exact, key mapping identical, checkpoint round trip lossless."""
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _stream(seed, n, lo, hi):
    rng = np.random.RandomState(seed)
    for _ in range(n):
        yield (rng.randn(int(rng.randint(lo, hi))) * 0.1).astype(np.float32)


def test_token_budget_batches_match_reference():
    from audio8_b200.feed import token_budget_batches
    g = np.load(os.path.join(GOLD, "feed.npz"))
    seed, n, lo, hi, max_len, budget = (int(v) for v in g["pre_cfg"])
    got = list(token_budget_batches((s[:max_len] for s in _stream(seed, n, lo, hi)), budget, max_len))
    assert len(got) == int(g["pre_n"][0])
    for i, b in enumerate(got):
        assert b.dtype == np.float32 and b.shape == g[f"pre_{i}"].shape and (b == g[f"pre_{i}"]).all(), f"batch {i}"
    # without the reference's accidents: no sample is lost and every batch starts from max_length again
    fixed = list(token_budget_batches((s[:max_len] for s in _stream(seed, n, lo, hi)), budget, max_len, reference_quirks=False))
    assert sum(b.shape[0] for b in fixed) > sum(b.shape[0] for b in got)
    assert all(b.shape[0] * b.shape[1] >= budget for b in fixed)


def test_collate_padded_matches_reference():
    from audio8_b200.feed import collate_padded
    g = np.load(os.path.join(GOLD, "feed.npz"))
    audios = list(_stream(9, 5, 2000, 6000))
    rng = np.random.RandomState(3)
    toks = [rng.randint(4, 32, size=int(rng.randint(3, 12))) for _ in range(5)]
    order = [3, 0, 4]
    sig, lens, ids, tl = collate_padded([audios[i] for i in order], [toks[i] for i in order], int(g["ft_pad"][0]), 12)
    assert (sig == g["ft_signal"]).all() and sig.dtype == np.float32
    assert (lens == g["ft_signal_lengths"]).all() and lens.dtype == g["ft_signal_lengths"].dtype
    assert (ids == g["ft_token_ids"]).all() and (tl == g["ft_token_lengths"]).all()
    with pytest.raises(ValueError):
        collate_padded(audios[:1], [np.arange(20)], 0, 12)


@pytest.mark.parametrize("name,ctc,sr", [("pretrain", False, 16), ("ctc16", True, 16), ("ctc8", True, 8)])
def test_fairseq_key_map_matches_reference(name, ctc, sr):
    from audio8_b200.wav2vec2 import fairseq_key_map
    want = json.load(open(os.path.join(GOLD, "fairseq_keymap.json")))[name]
    want = {k: v for k, v in want.items() if k != "untouched.key"}
    assert fairseq_key_map(3, ctc, sr) == want


@pytest.mark.parametrize("ctc", [False, True])
def test_fairseq_checkpoint_round_trip(ctc, tmp_path):
    """our state_dict -> fairseq key names -> file -> load_fairseq_bin: every tensor back in place, nothing missing"""
    from audio8_b200 import wav2vec2 as W
    torch.manual_seed(0)
    kw = dict(d_model=128, num_heads=2, num_layers=2, d_ff=256)
    src = W.create_acoustic_model(32, **kw) if ctc else W.create_model(final_dim=64, num_vq_vars=24, **kw)
    inv = {v: k for k, v in W.fairseq_key_map(2, ctc, 16).items()}
    fs = {inv.get(k, k): v.clone() for k, v in src.state_dict().items()}
    assert any(k.startswith("w2v_encoder.") for k in fs) == ctc
    path = str(tmp_path / "fairseq.pt")
    torch.save({"model": fs, "args": None}, path)
    torch.manual_seed(1)
    dst = W.create_acoustic_model(32, **kw) if ctc else W.create_model(final_dim=64, num_vq_vars=24, **kw)
    res = W.load_fairseq_bin(dst, path, ctc=ctc, sr=16)
    assert res == {"missing": [], "unexpected": []}
    a, b = src.state_dict(), dst.state_dict()
    assert set(a) == set(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # a checkpoint without one of the mapped keys fails like the reference (KeyError from the pop)
    del fs[next(iter(inv.values()))]
    torch.save({"model": fs}, path)
    with pytest.raises(KeyError):
        W.load_fairseq_bin(dst, path, ctc=ctc, sr=16)


def test_device_feed_order_and_values_cpu():
    from audio8_b200.feed import DeviceFeed
    batches = [np.random.RandomState(i).randn(3, 100 + i).astype(np.float32) for i in range(7)]
    got = list(DeviceFeed(iter(batches), "cpu", depth=2))
    assert len(got) == 7
    for a, b in zip(got, batches):
        assert (a.numpy() == b).all()


@pytest.mark.gpu
def test_device_feed_cuda():
    """pinned ring + copy stream: values, order, tuple batches, lengths -> pad mask on the device"""
    from audio8_b200.feed import DeviceFeed, collate_padded, pad_mask_from_lengths
    rng = np.random.RandomState(0)
    batches = []
    for i in range(9):
        audios = [rng.randn(rng.randint(500, 4000)).astype(np.float32) for _ in range(4)]
        toks = [rng.randint(4, 32, size=rng.randint(2, 9)) for _ in range(4)]
        batches.append(collate_padded(audios, toks, 1, 12))
    feed = DeviceFeed(iter(batches), "cuda", depth=2)
    n = 0
    for (sig, lens, ids, tl), ref in zip(feed, batches):
        assert sig.is_cuda and lens.dtype == torch.int32
        torch.cuda.synchronize()
        assert (sig.cpu().numpy() == ref[0]).all() and (ids.cpu().numpy() == ref[2]).all()
        pm = pad_mask_from_lengths(lens, sig.shape[1])
        assert (pm.sum(-1).cpu().numpy() == ref[1]).all()
        n += 1
    assert n == 9 and feed.h2d_bytes == sum(sum(a.nbytes for a in b) for b in batches)
