"""Shared GEMM test cases: each returns (spec, check) where check() compares the outputs against plain
PyTorch fp32 math on the same bf16-rounded inputs.  Used on CPU with tests/emu.py and on the GPU with the
tcgen05 kernel."""
import math

import torch
import torch.nn.functional as F

from audio8_b200 import gemm_specs as G
from audio8_b200.ops import ACT_GELU, ACT_GELU_DZ, ACT_NONE, AUX_ADD, AUX_MUL, AUX_MUL_GELU_GRAD, OUT_BF16, OUT_F32

from emu import gelu, gelu_grad


def _r(shape, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).to(dev)


def _cmp(got, want, tol, what):
    got, want = got.float().cpu(), want.float().cpu()
    err = (got - want).abs().max().item()
    scale = want.abs().max().item() + 1e-6
    assert err <= tol * scale, f"{what}: max abs err {err:.4g} vs scale {scale:.4g}"


def case_linear_fwd(dev, M=300, K=192, N=136, act=ACT_GELU, f32=False):
    x, w = _r((M, K), dev, 1), _r((N, K), dev, 2, 0.1)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3)).to(dev)
    aux = _r((M, N), dev, 4)
    out = torch.zeros(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)
    z = torch.zeros(M, N, dtype=torch.float16 if act == ACT_GELU_DZ else torch.bfloat16, device=dev)
    spec = G.linear_fwd(x, w, out, bias, act, z, aux, AUX_ADD, OUT_F32 if f32 else OUT_BF16)

    def check():
        zz = x.float() @ w.float().t() + bias
        _cmp(z, gelu_grad(zz) if act == ACT_GELU_DZ else zz, 1e-2, "linear z_out")
        _cmp(out, (gelu(zz) if act != ACT_NONE else zz) + aux.float(), 1e-2, "linear out")
    return spec, check


def case_linear_dgrad(dev, M=260, N=200, K=320, mode=AUX_MUL_GELU_GRAD, colsum=False):
    dy, w, z = _r((M, N), dev, 5), _r((N, K), dev, 6, 0.1), _r((M, K), dev, 7)
    if mode == AUX_MUL:  # the stored gelu' factors are fp16
        z = z.to(torch.float16)
    dx = torch.zeros(M, K, dtype=torch.bfloat16, device=dev)
    cs = torch.full((K,), 0.5, dtype=torch.float32, device=dev) if colsum else None  # accumulated INTO: starts non-zero
    spec = G.linear_dgrad(dy, w, dx, z, mode, colsum=cs)

    def check():
        f = gelu_grad(z.float()) if mode == AUX_MUL_GELU_GRAD else z.float()
        _cmp(dx, (dy.float() @ w.float()) * f, 1e-2, "linear dgrad")
        if colsum:  # the fused bias gradient: column sums of the output AS STORED (bf16), fp32 accumulation
            _cmp(cs, 0.5 + dx.float().sum(0), 1e-4, "linear dgrad column sums")
    return spec, check


def case_linear_wgrad(dev, M=1000, N=200, K=328):
    dy, x = _r((M, N), dev, 8), _r((M, K), dev, 9)
    dw = torch.zeros(N, K, dtype=torch.float32, device=dev)
    spec = G.linear_wgrad(dy, x, dw)

    def check():
        _cmp(dw, dy.float().t() @ x.float(), 2e-3, "linear wgrad")
    return spec, check


def _conv_pack(w):  # [Cout, Cin, k] -> [Cout, k*Cin]
    return w.permute(0, 2, 1).reshape(w.shape[0], -1).contiguous()


def case_conv_fwd(dev, B=2, Lin=301, C=128, Cout=192, k=3, s=2):
    Lout = (Lin - k) // s + 1
    x, w = _r((B, Lin, C), dev, 10), _r((Cout, C, k), dev, 11, 0.05)
    y = torch.zeros(B, Lout, Cout, dtype=torch.bfloat16, device=dev)
    z = torch.zeros(B, Lout, Cout, dtype=torch.float16, device=dev)
    spec = G.conv_fwd(x, _conv_pack(w), y, k, s, z)

    def check():
        zz = F.conv1d(x.float().cpu().transpose(1, 2), w.float().cpu(), stride=s).transpose(1, 2)
        _cmp(z, gelu_grad(zz), 1e-2, "conv z_out = gelu'(pre-activation)")
        _cmp(y, gelu(zz), 1e-2, "conv y")
    return spec, check


def conv_wt(w, s, p):  # [Cout, Cin, k] -> [Cin, ntaps*Cout] for phase p
    taps = G.conv_dgrad_taps(w.shape[2], s, p)
    return torch.cat([w[:, :, j].t() for j in taps], 1).contiguous()


def case_conv_dgrad(dev, B=2, Lin=301, C=128, Cout=192, k=3, s=2):
    Lout = (Lin - k) // s + 1
    dz, w, zprev = _r((B, Lout, Cout), dev, 12), _r((Cout, C, k), dev, 13, 0.05), _r((B, Lin, C), dev, 14).to(torch.float16)
    dx = torch.zeros(B, Lin, C, dtype=torch.bfloat16, device=dev)
    specs = [G.conv_dgrad(dz, conv_wt(w, s, p), dx, k, s, p, zprev) for p in range(s)]

    def check():
        g = F.conv_transpose1d(dz.float().cpu().transpose(1, 2), w.float().cpu(), stride=s)
        g = F.pad(g, (0, Lin - g.shape[-1])).transpose(1, 2)
        _cmp(dx, g * zprev.float().cpu(), 1e-2, "conv dgrad (* stored gelu')")
    return specs, check


def case_conv_wgrad(dev, B=3, Lin=411, C=128, Cout=192, k=3, s=2):
    Lout = (Lin - k) // s + 1
    dz, x = _r((B, Lout, Cout), dev, 15), _r((B, Lin, C), dev, 16)
    dwk = torch.zeros(Cout, k * C, dtype=torch.float32, device=dev)
    spec = G.conv_wgrad(dz, x, dwk, k, s)

    def check():
        xx = x.float().cpu().transpose(1, 2).requires_grad_(False)
        w = torch.zeros(Cout, C, k, requires_grad=True)
        F.conv1d(xx, w, stride=s).backward(dz.float().cpu().transpose(1, 2))
        _cmp(dwk, _conv_pack(w.grad), 2e-3, "conv wgrad")
    return spec, check


def posconv_pack(w, groups, transpose=False):
    """w [D, cg, k] -> [D, k*64] rows = output channel (or input channel if transpose), col = j*64 + c, zero pad"""
    D, cg, k = w.shape
    wg = w.view(groups, cg, cg, k)  # [g, co, ci, j]
    if transpose:
        wg = wg.permute(0, 2, 1, 3)  # [g, ci, co, j]
    out = torch.zeros(groups, cg, k, 64, dtype=w.dtype, device=w.device)
    out[..., :cg] = wg.permute(0, 1, 3, 2)
    return out.reshape(D, k * 64).contiguous()


def case_posconv(dev, B=2, T=70, D=128, groups=16, k=16):
    cg = D // groups
    pad_l = k // 2 - 1
    x, w = _r((B, T, D), dev, 17), _r((D, cg, k), dev, 18, 0.2)
    bias = torch.randn(D, generator=torch.Generator().manual_seed(19)).to(dev)
    out = torch.zeros(B, T, D, dtype=torch.bfloat16, device=dev)
    z = torch.zeros(B, T, D, dtype=torch.float16, device=dev)
    f = G.posconv_fwd(x, posconv_pack(w, groups), out, bias, groups, k, pad_l, z)
    dz, res = _r((B, T, D), dev, 20), _r((B, T, D), dev, 21)
    dx = torch.zeros_like(out)
    d = G.posconv_dgrad(dz, posconv_pack(w, groups, True), dx, groups, k, pad_l, res)
    dwp = torch.zeros(groups, k * 64, 64, dtype=torch.float32, device=dev)
    wg = G.posconv_wgrad(dz, x, dwp, groups, k, pad_l)

    def check():
        xc = x.float().cpu().transpose(1, 2).requires_grad_(True)
        wc = w.float().cpu().requires_grad_(True)
        zz = F.conv1d(F.pad(xc, (pad_l, k // 2)), wc, bias.cpu(), groups=groups)
        _cmp(z, gelu_grad(zz.detach()).transpose(1, 2), 1e-2, "posconv z_out = gelu'(pre-activation)")
        _cmp(out, (gelu(zz) + xc).transpose(1, 2), 1e-2, "posconv out")
        zz.backward(dz.float().cpu().transpose(1, 2))
        _cmp(dx, xc.grad.transpose(1, 2) + res.float().cpu(), 1e-2, "posconv dgrad")
        got = dwp.cpu().view(groups, k, 64, 64)[:, :, :cg, :cg].permute(0, 3, 2, 1).reshape(D, cg, k)  # [g,co,ci,j]
        _cmp(got, wc.grad, 2e-3, "posconv wgrad")
    return [f, d, wg], check


def case_attention(dev, B=2, T=150, H=2):
    D = 64 * H
    Tp = (T + 7) // 8 * 8
    qkv = _r((B, T, 3 * D), dev, 22)
    scale = 1.0 / math.sqrt(64)
    s = torch.zeros(B, H, T, Tp, dtype=torch.float32, device=dev)
    spec_s = G.attn_scores(qkv, s, H, scale)
    p = torch.zeros(B, H, T, Tp, dtype=torch.bfloat16, device=dev)
    p[..., :T] = torch.softmax(_r((B, H, T, T), dev, 23).float(), -1).to(torch.bfloat16)
    ctx = torch.zeros(B, T, D, dtype=torch.bfloat16, device=dev)
    spec_c = G.attn_context(p, qkv, ctx, H)
    dctx = _r((B, T, D), dev, 24)
    dp = torch.zeros_like(s)
    spec_dp = G.attn_dprobs(dctx, qkv, dp, H)
    ds = torch.zeros_like(p)
    ds[..., :T] = _r((B, H, T, T), dev, 25, 0.1)
    dqkv = torch.zeros_like(qkv)
    spec_dq = G.attn_dq(ds, qkv, dqkv, H, scale)
    spec_dk = G.attn_dk(ds, qkv, dqkv, H, scale)
    spec_dv = G.attn_dv(p, dctx, dqkv, H)

    def check():
        q, k, v = (qkv.float().cpu()[..., i * D:(i + 1) * D].view(B, T, H, 64).transpose(1, 2) for i in range(3))
        _cmp(s[..., :T], scale * q @ k.transpose(-1, -2), 1e-2, "attn scores")
        pc, dsc = p.float().cpu()[..., :T], ds.float().cpu()[..., :T]
        _cmp(ctx, (pc @ v).transpose(1, 2).reshape(B, T, D), 1e-2, "attn ctx")
        dc = dctx.float().cpu().view(B, T, H, 64).transpose(1, 2)
        _cmp(dp[..., :T], dc @ v.transpose(-1, -2), 1e-2, "attn dP")
        want = torch.cat([(scale * dsc @ k).transpose(1, 2).reshape(B, T, D),
                          (scale * dsc.transpose(-1, -2) @ q).transpose(1, 2).reshape(B, T, D),
                          (pc.transpose(-1, -2) @ dc).transpose(1, 2).reshape(B, T, D)], -1)
        _cmp(dqkv, want, 1e-2, "attn dqkv")
    return [spec_s, spec_c, spec_dp, spec_dq, spec_dk, spec_dv], check


ALL_CASES = dict(
    linear_fwd=case_linear_fwd,
    linear_fwd_f32=lambda dev: case_linear_fwd(dev, M=130, K=64, N=32, act=ACT_NONE, f32=True),
    linear_fwd_gelu_dz=lambda dev: case_linear_fwd(dev, act=ACT_GELU_DZ),
    linear_dgrad=case_linear_dgrad,
    linear_dgrad_mul=lambda dev: case_linear_dgrad(dev, mode=AUX_MUL),
    linear_dgrad_mul_colsum=lambda dev: case_linear_dgrad(dev, mode=AUX_MUL, colsum=True),
    linear_dgrad_mul_colsum_ffn=lambda dev: case_linear_dgrad(dev, M=1000, N=256, K=1024, mode=AUX_MUL, colsum=True),
    linear_wgrad=case_linear_wgrad,
    conv_fwd=case_conv_fwd,
    conv_fwd_k2=lambda dev: case_conv_fwd(dev, Lin=200, k=2, s=2),
    conv_dgrad=case_conv_dgrad,
    conv_dgrad_k2=lambda dev: case_conv_dgrad(dev, Lin=200, k=2, s=2),
    conv_wgrad=case_conv_wgrad,
    posconv=case_posconv,
    attention=case_attention,
)


def run_case(name, dev, backend):
    specs, check = ALL_CASES[name](dev)
    if not isinstance(specs, (list, tuple)):
        specs = [specs]
    for sp in specs:
        backend.gemm(sp)
    if dev != "cpu":
        torch.cuda.synchronize()
    check()
