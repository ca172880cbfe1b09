"""Per-kernel parity on the GPU: every row / index / loss kernel of the C ABI against the pure-PyTorch fp32
emulation of the same ABI (tests/emu.py), on seeded inputs.  bf16 outputs: max|err| <= 2e-2 * max|ref| (one
bf16 rounding of an fp32 value is 2^-9 relative; sums of a few of them stay well inside); fp32 outputs 1e-4."""
import math

import numpy as np
import pytest
import torch

import emu

pytestmark = pytest.mark.gpu


def _g(seed):
    return torch.Generator().manual_seed(seed)


def _bf(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=_g(seed)) * scale).to(torch.bfloat16)


def _close(got, want, tol, what):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item() + 1e-6
    assert err <= tol * scale, f"{what}: max abs err {err:.4g} vs scale {scale:.4g}"


@pytest.fixture(scope="module")
def be():
    from audio8_b200 import ops
    return ops.backend()


E = emu.EmuOps()


@pytest.mark.parametrize("R,C", [(37, 128), (300, 512), (1000, 768), (64, 1024)])
def test_layernorm(be, R, C):
    x, h = _bf((R, C), 1), _bf((R, C), 2)
    gamma = 1 + 0.1 * torch.randn(C, generator=_g(3))
    beta = 0.1 * torch.randn(C, generator=_g(4))
    dy = _bf((R, C), 5)
    dyf = torch.randn(R, C, generator=_g(6))
    for use_h in (False, True):
        ref = E.layernorm_fwd(x, gamma, beta, 1e-5, h=h if use_h else None, want_f32=True)
        got = be.layernorm_fwd(x.cuda(), gamma.cuda(), beta.cuda(), 1e-5, h=h.cuda() if use_h else None, want_f32=True)
        for i, name in enumerate(["y", "y_f32", "s", "mean", "rstd"]):
            _close(got[i], ref[i], 2e-2 if i < 3 else 1e-3, f"ln fwd {name} (h={use_h})")
        s, mean, rstd = ref[2], ref[3], ref[4]
        rb = E.layernorm_bwd(dy, s, mean, rstd, gamma, dy_f32=dyf, want_dh=use_h, want_dbias=True)
        gb = be.layernorm_bwd(dy.cuda(), s.cuda(), mean.cuda(), rstd.cuda(), gamma.cuda(), dy_f32=dyf.cuda(),
                              want_dh=use_h, want_dbias=True)
        for i, name in enumerate(["ds", "dh", "dgamma", "dbeta", "dbias"]):
            if rb[i] is None:
                assert gb[i] is None
                continue
            _close(gb[i], rb[i], 2e-2, f"ln bwd {name} (h={use_h})")


def test_layernorm_dropout_consistency(be):
    """forward and backward regenerate the same Philox mask; keep-rate matches p"""
    R, C, p = 256, 768, 0.1
    x, h = torch.zeros(R, C, dtype=torch.bfloat16).cuda(), _bf((R, C), 2).cuda()
    gamma, beta = torch.ones(C).cuda(), torch.zeros(C).cuda()
    y, _, s, mean, rstd = be.layernorm_fwd(x, gamma, beta, 1e-5, h=h, p_h=p, seed_h=1234)
    mask_fwd = s.float() != 0  # x == 0, so s = drop(h)
    keep = mask_fwd.float().mean().item()
    assert abs(keep - (1 - p)) < 0.01, keep
    ds, dh, *_ = be.layernorm_bwd(_bf((R, C), 3).cuda(), s, mean, rstd, gamma, want_dh=True, p_h=p, seed_h=1234)
    live = ds.float() != 0
    assert ((dh.float() != 0) == mask_fwd)[live].float().mean().item() > 0.9999
    nz = dh.float() != 0
    _close(dh.float()[nz], (ds.float() / (1 - p))[nz], 1e-2, "dh scale")


def test_colsum_gelu_dropout_cast(be):
    x = _bf((1000, 3072), 1)
    _close(be.colsum(x.cuda()), E.colsum(x), 2e-3, "colsum")
    dy, z = _bf((300, 512), 2), _bf((300, 512), 3, 2.0)
    _close(be.gelu_bwd(dy.cuda(), z.cuda()), E.gelu_bwd(dy, z), 2e-2, "gelu_bwd")
    gz = torch.rand(300, 512, generator=_g(5)).to(torch.float16)
    _close(be.mul_dgelu(dy.cuda(), gz.cuda()), E.mul_dgelu(dy, gz), 1e-2, "mul_dgelu")
    for t in (x, torch.randn(64, 512, generator=_g(4))):
        d = be.dropout(t.cuda(), 0.1, 99)
        kept = (d.float() != 0).float().mean().item()
        assert abs(kept - 0.9) < 0.01
        d2 = be.dropout(t.cuda(), 0.1, 99)
        assert torch.equal(d, d2)
        nz = d.float() != 0
        _close(d.float()[nz], (t.cuda().float() / 0.9)[nz], 1e-2, "dropout scale")
    f = torch.randn(513, 40, generator=_g(5))
    assert torch.equal(be.cast(f.cuda(), torch.bfloat16).cpu(), f.to(torch.bfloat16))
    for side in (False, True):
        assert torch.equal(be.split3(f.cuda(), side).cpu(), E.split3(f, side))


def test_log_softmax(be):
    x = torch.randn(3, 50, 32, generator=_g(1)) * 3
    y = be.log_softmax_fwd(x.cuda())
    _close(y, torch.log_softmax(x, -1), 1e-5, "log_softmax fwd")
    dy_tbv = torch.randn(50, 3, 32, generator=_g(2))  # CTC hands back [T,B,V]
    got = be.log_softmax_bwd(dy_tbv.cuda().transpose(0, 1), y)
    _close(got, E.log_softmax_bwd(dy_tbv.transpose(0, 1), torch.log_softmax(x, -1)), 2e-2, "log_softmax bwd")


@pytest.mark.parametrize("B,L", [(2, 4000), (3, 16000)])
def test_conv0(be, B, L):
    C, k, s = 512, 10, 5
    x = torch.randn(B, L, generator=_g(1)) * 0.1
    w = (torch.rand(C, k, generator=_g(2)) * 2 - 1) * math.sqrt(3.0 / k)
    gamma = 1 + 0.1 * torch.randn(C, generator=_g(3))
    beta = 0.1 * torch.randn(C, generator=_g(4))
    mean_r, rstd_r, _ = E.conv0_stats(x, w, k, s, 1e-5)
    mean, rstd, mom = be.conv0_stats(x.cuda(), w.cuda(), k, s, 1e-5)
    _close(mean, mean_r, 1e-4, "conv0 mean")
    _close(rstd, rstd_r, 1e-4, "conv0 rstd")
    y_r = E.conv0_fwd(x, w, gamma, beta, mean_r, rstd_r, k, s)
    y = be.conv0_fwd(x.cuda(), w.cuda(), gamma.cuda(), beta.cuda(), mean, rstd, k, s)
    _close(y, y_r, 1e-2, "conv0 fwd")
    da = _bf(tuple(y_r.shape), 5)
    ref = E.conv0_bwd(x, w, gamma, beta, mean_r, rstd_r, None, k, s, da)
    got = be.conv0_bwd(x.cuda(), w.cuda(), gamma.cuda(), beta.cuda(), mean, rstd, mom, k, s, da.cuda())
    for g, r, name in zip(got, ref, ["dw", "dgamma", "dbeta"]):
        _close(g, r, 5e-3, "conv0 bwd " + name)


def test_rows_and_masks(be):
    src = torch.randn(200, 96, generator=_g(1))
    idx = torch.randperm(200, generator=_g(2))[:57].sort().values.int()
    for dt in (torch.float32, torch.bfloat16):
        assert torch.equal(be.rows_gather(src.cuda(), idx.cuda(), dt).cpu(), E.rows_gather(src, idx, dt))
        rows = torch.randn(57, 96, generator=_g(3))
        assert torch.equal(be.rows_scatter(rows.cuda(), idx.cuda(), 200, dt).cpu(), E.rows_scatter(rows, idx, 200, dt))
    xb = _bf((200, 96), 4)
    vec = torch.randn(96, generator=_g(5))
    a, b = xb.clone().cuda(), xb.clone()
    be.rows_set(a, idx.cuda(), vec.cuda())
    E.rows_set(b, idx, vec)
    assert torch.equal(a.cpu(), b)
    a, b = xb.clone().cuda(), xb.clone()
    _close(be.rows_set_bwd(a, idx.cuda()), E.rows_set_bwd(b, idx), 1e-3, "rows_set_bwd dvec")
    assert torch.equal(a.cpu(), b)
    x3 = _bf((4, 50, 96), 6)
    rk = (torch.rand(4, 50, generator=_g(7)) > 0.3).to(torch.uint8)
    cz = (torch.rand(4, 96, generator=_g(8)) > 0.8).to(torch.uint8)
    for r_, c_ in ((rk, None), (None, cz), (rk, cz)):
        a, b = x3.clone().cuda(), x3.clone()
        be.mask_apply(a, r_.cuda() if r_ is not None else None, c_.cuda() if c_ is not None else None)
        E.mask_apply(b, r_, c_)
        assert torch.equal(a.cpu(), b)


@pytest.mark.parametrize("train", [True, False])
def test_vq(be, train):
    R, G, V, vd = 130, 2, 320, 128
    z = torch.randn(R, G * V, generator=_g(1)) * 4
    noise = -torch.empty(R * G, V).exponential_(generator=_g(2)).log() if train else None
    vars2d = torch.rand(G * V, vd, generator=_g(3))
    ref = E.vq_fwd(z, noise, 0.5, vars2d, G)
    got = be.vq_fwd(z.cuda(), noise.cuda() if train else None, 0.5, vars2d.cuda(), G)
    assert torch.equal(got[2].cpu(), ref[2]), "VQ arg-max indices must be bit-exact"
    assert torch.equal(got[0].cpu(), ref[0]), "selected codewords must be an exact gather"
    _close(got[3], ref[3], 1e-4, "avg_sums")
    _close(got[4], ref[4], 1e-4, "ppl")
    dq = torch.randn(R, G * vd, generator=_g(4))
    a_dot = torch.einsum("rgd,gvd->rgv", dq.view(R, G, vd), vars2d.view(G, V, vd)).reshape(R, G * V)
    dppl = torch.tensor(-10.0 / 640)
    rb = E.vq_bwd(z, noise, 0.5, G, vd, a_dot, dq, ref[2], ref[3], ref[4], dppl)
    gb = be.vq_bwd(z.cuda(), noise.cuda() if train else None, 0.5, G, vd, a_dot.cuda(), dq.cuda(), got[2], got[3],
                   got[4], dppl.cuda())
    _close(gb[0], rb[0], 2e-2, "vq dz")
    _close(gb[1], rb[1], 1e-4, "vq dvars")


def test_contrastive(be):
    B, Tm, C, K = 3, 40, 256, 100
    R = B * Tm
    x, y = torch.randn(R, C, generator=_g(1)), torch.randn(R, C, generator=_g(2))
    rng = np.random.RandomState(5)
    own = np.repeat(np.arange(Tm), K)[None, :]
    idx = rng.randint(0, Tm - 1, (B, K * Tm))
    idx = np.where(idx >= own, idx + 1, idx) + (np.arange(B) * Tm)[:, None]
    idx = torch.from_numpy(idx.astype(np.int32))
    ppl = torch.tensor(123.4)
    l_r, ce_r, sv_r = E.contrastive_fwd(x, y, idx, ppl, 640.0, 0.1, 10.0)
    l_g, ce_g, sv_g = be.contrastive_fwd(x.cuda(), y.cuda(), idx.cuda(), ppl.cuda(), 640.0, 0.1, 10.0)
    assert abs(l_g.item() - l_r.item()) < 1e-5 * abs(l_r.item()) and abs(ce_g.item() - ce_r.item()) < 1e-5 * ce_r.item()
    dce = torch.tensor(0.1)
    dx_r, dy_r = E.contrastive_bwd(x, y, idx, sv_r, dce)
    dx_g, dy_g = be.contrastive_bwd(x.cuda(), y.cuda(), idx.cuda(), sv_g, dce.cuda())
    _close(dx_g, dx_r, 1e-4, "contrastive dx")
    _close(dy_g, dy_r, 1e-4, "contrastive dy")


def test_weight_prep(be):
    """parameter re-layout kernels (csrc/wprep.cu) against the emulation"""
    w = torch.randn(192, 128, 3, generator=_g(1)) * 0.1
    for want_t in (False, True):
        wk, wts = be.conv_pack(w.cuda(), 2, want_t)
        wk_r, wts_r = E.conv_pack(w, 2, want_t)
        assert torch.equal(wk.cpu(), wk_r)
        if want_t:
            for a, b in zip(wts, wts_r):
                assert torch.equal(a.cpu(), b)
    w2 = torch.randn(64, 128, 2, generator=_g(2))
    wk, wts = be.conv_pack(w2.cuda(), 2, True)
    wk_r, wts_r = E.conv_pack(w2, 2, True)
    assert torch.equal(wk.cpu(), wk_r) and all(torch.equal(a.cpu(), b) for a, b in zip(wts, wts_r))
    dwk = torch.randn(192, 3 * 128, generator=_g(3))
    assert torch.equal(be.conv_unpack(dwk.cuda(), 128, 3).cpu(), E.conv_unpack(dwk, 128, 3))
    D, cg, k = 768, 48, 128
    v = torch.randn(D, cg, k, generator=_g(4)) * 0.02
    g = 0.5 + torch.rand(1, 1, k, generator=_g(5))
    wp, wpt, n2 = be.posconv_pack(g.cuda(), v.cuda(), True)
    wp_r, wpt_r, n2_r = E.posconv_pack(g, v, True)
    _close(n2, n2_r, 1e-5, "posconv norm2")
    _close(wp, wp_r, 1e-2, "posconv wp")
    _close(wpt, wpt_r, 1e-2, "posconv wpt")
    dwp = torch.randn(D // cg, k * 64, 64, generator=_g(6))
    dv, dg = be.posconv_wn_bwd(dwp.cuda(), g.cuda(), v.cuda(), n2)
    dv_r, dg_r = E.posconv_wn_bwd(dwp, g, v, n2_r)
    _close(dv, dv_r, 1e-4, "posconv dv")
    _close(dg, dg_r, 1e-4, "posconv dg")
    srcs = [torch.randn(n, generator=_g(10 + i)) for i, n in enumerate((768 * 768, 17, 3072 * 768, 768))]
    dsts = [torch.zeros(768 * 768, dtype=torch.bfloat16), torch.zeros(17, dtype=torch.bfloat16),
            torch.zeros(3072 * 768, dtype=torch.bfloat16), torch.zeros(768)]
    pairs = [(s_.cuda(), d_.cuda()) for s_, d_ in zip(srcs, dsts)]
    cache = {}
    be.cast_multi(pairs, cache)
    be.cast_multi(pairs, cache)
    for (s_, d_) in pairs:
        assert torch.equal(d_.cpu(), s_.cpu().to(d_.dtype))


@pytest.mark.parametrize("B,H,T,mask,pdrop", [(2, 2, 100, False, 0.0), (1, 3, 128, False, 0.0), (2, 2, 333, True, 0.0),
                                              (1, 2, 749, False, 0.0), (2, 2, 300, True, 0.1), (1, 2, 749, False, 0.1)])
def test_fused_attention(be, B, H, T, mask, pdrop):
    """fused tcgen05 attention (csrc/attn.cu) vs fp32 softmax attention on the same bf16 inputs; with dropout the
    reference uses the keep mask exported by a8_attn_dropmask (the function of (seed, b, h, q, k) the kernels use)"""
    D = H * 64
    qkv = _bf((B, T, 3 * D), 21, 1.0)
    qkv[..., :D] *= 1.5  # scores with a spread of a few units
    dctx = _bf((B, T, D), 22)
    keep_keys = None
    if mask:
        lens = [T - 37 * (b + 1) for b in range(B)]
        keep_keys = (torch.arange(T)[None, :] < torch.tensor(lens)[:, None]).to(torch.uint8)
    scale, seed = 0.125, 1234567
    kk = keep_keys.cuda() if mask else None
    ctx, lse = be.attn_fwd(qkv.cuda(), H, scale, kk, pdrop, seed)
    keep = be.attn_dropmask(B, H, T, pdrop, seed, "cuda").cpu() if pdrop > 0 else None
    if keep is not None:
        rate = 1.0 - keep.float().mean().item()
        assert abs(rate - pdrop) < 0.01, f"dropout rate {rate}"
    ref_ctx, _ = E.attn_fwd(qkv, H, scale, keep_keys, pdrop, seed, keep=keep)
    _close(ctx, ref_ctx, 2e-2, "attention ctx")
    # log2-sum-exp2 of the scaled scores
    q, k, v, p = E._attn_probs(qkv, H, scale, keep_keys)
    s = (q @ k.transpose(-1, -2)) * scale
    if mask:
        s = s.masked_fill(keep_keys[:, None, None, :] == 0, float("-inf"))
    _close(lse, torch.logsumexp(s, -1) / math.log(2.0), 1e-3, "attention lse")
    dbias = torch.full((3 * D,), 0.25, device="cuda")  # accumulated into: starts non-zero
    dqkv = be.attn_bwd(qkv.cuda(), ctx, dctx.cuda(), lse, H, scale, kk, pdrop, seed, dbias=dbias)
    ref_d = E.attn_bwd(qkv, ref_ctx, dctx, None, H, scale, keep_keys, pdrop, seed, keep=keep)
    for i, name in enumerate(["dQ", "dK", "dV"]):
        _close(dqkv[..., i * D:(i + 1) * D], ref_d[..., i * D:(i + 1) * D], 3e-2, "attention " + name)
    # the fused QKV bias gradient: column sums of dqkv exactly as stored (fp32 accumulation order aside)
    want_b = 0.25 + dqkv.float().sum((0, 1))
    err = (dbias - want_b).abs().max().item()
    assert err <= 1e-3 * (want_b.abs().max().item() + 1.0), f"attention dbias: max abs err {err:.4g}"


@pytest.mark.parametrize("qscale", [0.3, 1.5, 3.0])
def test_fused_attention_rel_l2_and_determinism(be, qscale):
    """The max-abs tolerance of test_fused_attention is blind to corrupted probabilities (flat softmax rows average them
    away): hold the kernels to the rel-L2 of a float64 softmax attention on PEAKED score distributions, within 2x of bf16
    rounding of the exact result, and to bit-identical results across launches (a TMEM buffer race shows up as both)."""
    B, H, T = 2, 4, 749
    D = H * 64
    qkv = _bf((B, T, 3 * D), 31, 1.0)
    qkv[..., :D] *= qscale
    dctx = _bf((B, T, D), 32)
    scale = 0.125
    x = qkv.double().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    p = torch.softmax(x[0] @ x[1].transpose(-1, -2) * scale, -1)
    o = (p @ x[2]).permute(0, 2, 1, 3).reshape(B, T, D)
    o.backward(dctx.double())
    want_d = x.grad.permute(1, 3, 0, 2, 4).reshape(B, T, 3 * D)

    def rel(a, b):
        return ((a.double().cpu() - b).norm() / b.norm()).item()

    first = None
    for rep in range(4):
        ctx, lse = be.attn_fwd(qkv.cuda(), H, scale, None, 0.0, 1)
        dqkv = be.attn_bwd(qkv.cuda(), ctx, dctx.cuda(), lse, H, scale, None, 0.0, 1)
        if first is None:
            first = (ctx.clone(), dqkv.clone())
            floor = rel(o.detach().bfloat16(), o.detach())
            assert rel(ctx, o.detach()) <= 2.0 * floor, f"ctx rel-L2 {rel(ctx, o.detach()):.5f} vs bf16 rounding {floor:.5f}"
            for i, name in enumerate(["dQ", "dK", "dV"]):
                e = rel(dqkv[..., i * D:(i + 1) * D], want_d[..., i * D:(i + 1) * D])
                assert e <= 6e-3, f"{name} rel-L2 {e:.5f}"
        else:
            assert torch.equal(ctx, first[0]) and torch.equal(dqkv, first[1]), f"launch {rep} differs from launch 0"


# ---- device-side draws (csrc/draws.cu): bit for bit against the emulation, then the distributions
# against the reference's numpy draws
def _rmax(B, T, p, L):
    return B * min(T, (int(p * T / float(L)) + 1) * L)


@pytest.mark.parametrize("B,T,p,L", [(6, 749, 0.65, 10), (1, 49, 0.65, 10), (3, 30, 0.65, 10), (40, 99, 0.65, 10),
                                     (8, 1499, 0.5, 4), (2, 12, 0.65, 10), (4, 20, 0.05, 10)])
def test_span_mask_draw_matches_emulation(be, B, T, p, L):
    import model_cases
    R_max = _rmax(B, T, p, L) + 3  # slack above the worst case is legal
    for seed in (1, 0x9E3779B97F4A7C15, 77):
        sd = torch.tensor([(seed * 31 + 5) & 0x3FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda")
        rows, mask = be.span_mask_draw(seed, sd, B, T, p, L, R_max, "cuda")
        rows_e, mask_e = E.span_mask_draw(seed, sd.cpu(), B, T, p, L, R_max, "cpu")
        assert torch.equal(rows.cpu(), rows_e), f"rows differ (seed {seed})"
        assert torch.equal(mask.cpu(), mask_e), f"mask differs (seed {seed})"
        if int(rows_e[R_max]) > 0:
            model_cases.check_span_mask(rows_e.numpy(), mask_e.numpy(), B, T, p, L, R_max)
        K = 7
        neg = be.negatives_draw(seed + 1, sd, rows, B, K)
        neg_e = E.negatives_draw(seed + 1, sd.cpu(), rows_e, B, K)
        assert torch.equal(neg.cpu(), neg_e), f"negatives differ (seed {seed})"
        if int(rows_e[R_max]) // B > 1:
            model_cases.check_negatives(neg_e.numpy(), int(rows_e[R_max]), B, K, R_max)


def test_span_mask_draw_rejects_impossible_shapes(be):
    from audio8_b200._lib import A8Error
    with pytest.raises(A8Error):  # R_max below the worst case
        be.span_mask_draw(1, None, 4, 200, 0.65, 10, 10, "cuda")
    with pytest.raises(A8Error):  # the spans cannot start at distinct frames
        be.span_mask_draw(1, None, 1, 3, 5.0, 1, 64, "cuda")


def test_device_draw_distributions_match_numpy_reference(be):
    """the device mode draws from the reference's distributions: per-frame masking frequency and masked-frame count of
    `create_mask` (numpy, reference order) vs the kernel over many seeds; negatives uniform over the other steps"""
    from audio8_b200.wav2vec2 import create_mask
    B, T, p, L, N = 4, 149, 0.65, 10, 400
    R_max = _rmax(B, T, p, L)
    np.random.seed(0)
    f_np = np.zeros(T)
    c_np = []
    for _ in range(N):
        m = create_mask((B, T), p, L)
        f_np += m.mean(0)
        c_np.append(m[0].sum())
    f_dev = np.zeros(T)
    c_dev = []
    for s in range(N):
        rows, mask = be.span_mask_draw(1000 + s, None, B, T, p, L, R_max, "cuda")
        m = mask.cpu().numpy()
        f_dev += m.mean(0)
        c_dev.append(m[0].sum())
    f_np /= N
    f_dev /= N
    # N*B = 1600 Bernoulli samples per frame: sigma <= 0.0125; spans make neighbouring frames correlated, not a frame's mean
    assert np.abs(f_np - f_dev).max() < 0.08, np.abs(f_np - f_dev).max()  # ~4.5 sigma of a difference
    assert abs(f_np.mean() - f_dev.mean()) < 0.01, (f_np.mean(), f_dev.mean())
    assert abs(np.mean(c_np) - np.mean(c_dev)) < 0.03 * np.mean(c_np), (np.mean(c_np), np.mean(c_dev))
    # the first frames are masked less often (no span can start before frame 0): the edge profile must agree too
    assert abs(f_np[:L].mean() - f_dev[:L].mean()) < 0.03
    # negatives: every other step of the utterance equally likely
    rows, _ = be.span_mask_draw(5, None, 2, 60, p, L, _rmax(2, 60, p, L), "cuda")
    n = int(rows[-1])
    Tm, K = n // 2, 20000
    neg = be.negatives_draw(9, None, rows, 2, K).cpu().numpy().reshape(-1, K)
    hist = np.bincount(neg[3], minlength=n)[:Tm]  # step 3 of utterance 0
    assert hist[3] == 0 and hist.sum() == K
    others = np.delete(hist, 3)
    expect = K / (Tm - 1)
    assert np.abs(others - expect).max() < 6 * np.sqrt(expect), (others.min(), others.max(), expect)
