"""Device-side draws (csrc/draws.cu) on the CPU: the algorithm the kernels implement is specified by
`tests/emu.py` (Philox4x32-10 counters, Floyd subsets, multiply-high range reduction) and the kernels are held to it bit for
bit on the GPU (tests/test_kernels.py).  Here the specification itself is checked: the generator against the published
Random123 known-answer vectors, the draws structurally, and their distributions against the reference's numpy
`create_mask` / `Sampler` (reference wav2vec2.py:189-216, 955-976)."""
import numpy as np

import emu
import model_cases


def _rmax(B, T, p, L):
    return B * min(T, (int(p * T / float(L)) + 1) * L)


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors: philox4x32 10 rounds"""
    got = [int(v) for v in emu.philox4x32_10(0, 0, 0, 0, 0)]
    assert got == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    got = [int(v) for v in emu.philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF)]
    assert got == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    got = [int(v) for v in emu.philox4x32_10(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, (0x299F31D0 << 32) | 0xA4093822)]
    assert got == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_span_mask_and_negatives_structure_cpu():
    for (B, T, p, L) in [(6, 749, 0.65, 10), (1, 49, 0.65, 10), (3, 30, 0.65, 10), (40, 99, 0.65, 10), (8, 1499, 0.5, 4),
                         (2, 12, 0.65, 10)]:
        R_max = _rmax(B, T, p, L)
        for seed in (1, 0x9E3779B97F4A7C15, 77):
            rows, mask = emu.emu_span_mask(seed, B, T, p, L, R_max)
            n = int(rows[R_max])
            if n:
                model_cases.check_span_mask(rows, mask, B, T, p, L, R_max)
            if n // B > 1:
                model_cases.check_negatives(emu.emu_negatives(seed + 1, n, B, 7, R_max), n, B, 7, R_max)
    # the same seed draws the same mask, another seed another one
    a = emu.emu_span_mask(5, 4, 149, 0.65, 10, _rmax(4, 149, 0.65, 10))
    b = emu.emu_span_mask(5, 4, 149, 0.65, 10, _rmax(4, 149, 0.65, 10))
    c = emu.emu_span_mask(6, 4, 149, 0.65, 10, _rmax(4, 149, 0.65, 10))
    assert np.array_equal(a[0], b[0]) and not np.array_equal(a[0], c[0])


def test_device_draw_distributions_match_the_reference_numpy_draws_cpu():
    from audio8_b200.wav2vec2 import create_mask
    B, T, p, L, N = 4, 149, 0.65, 10, 300
    R_max = _rmax(B, T, p, L)
    np.random.seed(0)
    f_np, c_np = np.zeros(T), []
    for _ in range(N):
        m = create_mask((B, T), p, L)
        f_np += m.mean(0)
        c_np.append(m[0].sum())
    f_dev, c_dev = np.zeros(T), []
    for s in range(N):
        _, m = emu.emu_span_mask(1000 + s, B, T, p, L, R_max)
        f_dev += m.mean(0)
        c_dev.append(m[0].sum())
    f_np /= N
    f_dev /= N
    # N*B = 1200 Bernoulli samples per frame: sigma of a difference ~0.02
    assert np.abs(f_np - f_dev).max() < 0.09, np.abs(f_np - f_dev).max()
    assert abs(f_np.mean() - f_dev.mean()) < 0.012, (f_np.mean(), f_dev.mean())
    assert abs(np.mean(c_np) - np.mean(c_dev)) < 0.03 * np.mean(c_np), (np.mean(c_np), np.mean(c_dev))
    assert abs(f_np[:L].mean() - f_dev[:L].mean()) < 0.035  # the edge profile (no span starts before frame 0)
    # negatives: uniform over the other masked steps of the utterance
    rows, _ = emu.emu_span_mask(5, 2, 60, p, L, _rmax(2, 60, p, L))
    n = int(rows[-1])
    Tm, K = n // 2, 20000
    neg = emu.emu_negatives(9, n, 2, K, _rmax(2, 60, p, L)).reshape(-1, K)
    hist = np.bincount(neg[3], minlength=n)[:Tm]
    assert hist[3] == 0 and hist.sum() == K
    expect = K / (Tm - 1)
    assert np.abs(np.delete(hist, 3) - expect).max() < 6 * np.sqrt(expect)
