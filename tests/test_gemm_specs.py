"""GEMM descriptors (conv windows, tap shifts, groups, heads) checked on CPU through the ABI emulation, and on
the GPU through the tcgen05 kernel, against plain PyTorch fp32 math."""
import pytest
import torch

import gemm_cases


@pytest.mark.parametrize("name", sorted(gemm_cases.ALL_CASES))
def test_spec_semantics_cpu(name, emu_backend):
    gemm_cases.run_case(name, "cpu", emu_backend)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(gemm_cases.ALL_CASES))
def test_tcgen05_gemm(name):
    from audio8_b200 import ops
    gemm_cases.run_case(name, "cuda", ops.backend())


@pytest.mark.gpu
@pytest.mark.parametrize("bn", [128, 256])
@pytest.mark.parametrize("name", ["linear_fwd", "linear_dgrad", "linear_wgrad", "conv_fwd", "conv_dgrad", "conv_wgrad"])
def test_tcgen05_gemm_cta_pair(name, bn):
    """the same cases with the 2-CTA (cta_group::2, 256 x BN pair tile) variant forced for every launch that allows it"""
    from audio8_b200 import gemm_specs as G
    from audio8_b200 import ops
    G.clear_spec_caches()
    G.FORCE.update(bn=bn, cluster=2, split=1)
    try:
        gemm_cases.run_case(name, "cuda", ops.backend())
    finally:
        G.FORCE.clear()
        G.clear_spec_caches()


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(4494, 2304, 768), (4494, 768, 3072), (1153, 640, 512), (257, 256, 64)])
@pytest.mark.parametrize("cluster", [1, 2])
def test_tcgen05_linear_large(M, N, K, cluster):
    """transformer-sized linears (odd number of 128-row tiles: the pair's second CTA runs past M) vs fp32 matmul"""
    from audio8_b200 import gemm_specs as G
    from audio8_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    x = (torch.randn(M, K, generator=g)).to(torch.bfloat16).cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).cuda()
    dy = (torch.randn(M, N, generator=g)).to(torch.bfloat16).cuda()
    G.clear_spec_caches()
    G.FORCE.update(cluster=cluster)
    try:
        out = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
        ops.backend().gemm(G.linear_fwd.raw(x, w, out, b))
        dx = torch.zeros(M, K, dtype=torch.bfloat16, device="cuda")
        ops.backend().gemm(G.linear_dgrad.raw(dy, w, dx))
        dw = torch.zeros(N, K, dtype=torch.float32, device="cuda")
        ops.backend().gemm(G.linear_wgrad.raw(dy, x, dw))
        torch.cuda.synchronize()
    finally:
        G.FORCE.clear()
        G.clear_spec_caches()
    ref = x.float() @ w.float().t() + b
    assert (out.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    ref = dy.float() @ w.float()
    assert (dx.float() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()
    ref = dy.float().t() @ x.float()
    assert (dw - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


def _group_case(dev, be, shapes):
    from audio8_b200 import gemm_specs as G
    g = torch.Generator().manual_seed(17)
    probs, specs = [], []
    for (M, N, K) in shapes:
        dy = torch.randn(M, N, generator=g).to(torch.bfloat16).to(dev)
        x = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
        dw = torch.full((N, K), float("nan"), dtype=torch.float32, device=dev)  # plain stores: no zero fill needed
        probs.append((dy, x, dw))
        specs.append(G.linear_wgrad_grouped(dy, x, dw))
    be.gemm_group(specs)
    for i, (dy, x, dw) in enumerate(probs):
        ref = dy.float().t() @ x.float()
        err = (dw - ref).abs().max().item()
        assert err <= 2e-3 * ref.abs().max().item() + 1e-6, f"group problem {i} {tuple(dw.shape)}: {err:.4g}"


def test_grouped_wgrad_cpu(emu_backend):
    _group_case("cpu", emu_backend, [(200, 136, 72), (200, 64, 264)])


@pytest.mark.gpu
@pytest.mark.parametrize("shapes", [
    [(4494, 2304, 768), (4494, 768, 768), (4494, 3072, 768), (4494, 768, 3072)] * 3,  # 12 problems, 324 tile pairs
    [(301, 384, 128), (301, 128, 128), (301, 256, 128), (301, 128, 256), (77, 136, 72)],  # ragged tiny shapes
    [(1000, 264, 520)] * 50,  # more than one launch (48 per group)
])
def test_grouped_wgrad_tcgen05(shapes):
    """a8_gemm_group: several weight-gradient problems with their own operands / extents in one persistent launch"""
    from audio8_b200 import ops
    _group_case("cuda", ops.backend(), shapes)


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,D,k", [(2, 749, 768, 128), (1, 300, 1024, 128), (3, 130, 256, 128), (2, 257, 768, 64)])
@pytest.mark.parametrize("window", [True, False])
def test_tcgen05_posconv_shapes(B, T, D, k, window):
    """positional conv fwd / dgrad / wgrad at the model's shapes (base: 16 x 48 channels, large: 16 x 64, C1: 16 x 16),
    through the tap-window kernel (gemm_tc_window.cu) and through the plain kernel on the same descriptors"""
    from audio8_b200 import gemm_specs as G
    from audio8_b200 import ops
    G.clear_spec_caches()
    G.FORCE.update(window=window)
    try:
        specs, check = gemm_cases.case_posconv("cuda", B=B, T=T, D=D, groups=16, k=k)
        assert (specs[0].spec().window_k16 is not None) == window
        for sp in specs:
            ops.backend().gemm(sp)
        torch.cuda.synchronize()
        check()
    finally:
        G.FORCE.clear()
        G.clear_spec_caches()
