"""CTC loss: oracle pinned on CPU against the fixtures generated from the unmodified reference; the CUDA kernels
checked on the GPU against the oracle (installed torch F.ctc_loss on CPU — the function the reference calls)
and against the same fixtures.  Tolerances (fp32 kernels): loss rtol 2e-5; gradients (occupancy probabilities
in [0,1]) atol 5e-5 + 2e-5*sqrt(T): alpha/beta are fp32 log-space values of magnitude ~|nll| (hundreds), so
each of the T recursion steps rounds at ~1.5e-5 and the error random-walks — the same holds for ATen's own
fp32 CPU/CUDA kernels against each other."""
import math
import os

import numpy as np
import pytest
import torch

import ref_ctc

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ctc_cases.npz"))
CASES = ["small", "rep", "infeasible", "mid"]


def _case(name):
    return (torch.from_numpy(GOLD[f"{name}_lp"]), torch.from_numpy(GOLD[f"{name}_targets"]),
            torch.from_numpy(GOLD[f"{name}_il"]), torch.from_numpy(GOLD[f"{name}_tl"]))


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("red", ["sum", "mean"])
def test_oracle_matches_reference_fixture(name, red):
    lp, tg, il, tl = _case(name)
    lp = lp.clone().requires_grad_(True)
    loss = ref_ctc.ctc_loss_reference(lp, il, tg, tl, 0, 1, 2, red)
    loss.backward()
    assert abs(loss.item() - GOLD[f"{name}_{red}_loss"][0]) <= 1e-5 * max(1, abs(loss.item()))
    np.testing.assert_allclose(lp.grad.numpy(), GOLD[f"{name}_{red}_grad"], atol=1e-6)


def test_numpy_restatement_matches_torch():
    lp, tg, il, tl = _case("rep")
    labels = [tg[b, : tl[b]].numpy() for b in range(tg.shape[0])]
    nll, grad = ref_ctc.ctc_numpy(lp.double().numpy(), il.numpy(), labels, 0)
    assert abs(nll.sum() - GOLD["rep_sum_loss"][0]) < 1e-4
    np.testing.assert_allclose(grad, GOLD["rep_sum_grad"], atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("red", ["sum", "mean"])
def test_cuda_ctc_matches_fixture(name, red):
    from audio8_b200.ctc import ctc_loss
    lp, tg, il, tl = _case(name)
    lpd = lp.cuda().requires_grad_(True)
    loss = ctc_loss(lpd, il.cuda(), tg.cuda(), tl, blank=0, pad=1, eos=2, reduction=red)
    loss.backward()
    want = GOLD[f"{name}_{red}_loss"][0]
    assert abs(loss.item() - want) <= 2e-5 * max(1.0, abs(want)), (loss.item(), want)
    np.testing.assert_allclose(lpd.grad.cpu().numpy(), GOLD[f"{name}_{red}_grad"], rtol=0,
                               atol=5e-5 + 2e-5 * math.sqrt(lp.shape[0]))


@pytest.mark.gpu
@pytest.mark.parametrize("T,B,S,V", [(50, 8, 10, 32), (250, 32, 50, 32), (750, 64, 150, 32), (300, 5, 40, 100),
                                     (1500, 16, 300, 32), (64, 3, 0, 8), (1500, 256, 300, 32)])
def test_cuda_ctc_matches_oracle_sweep(T, B, S, V):
    """config-5 sweep shapes; log-probs arrive as the transposed [B,T,V] view the trainer passes (train.py:39)"""
    from audio8_b200.ctc import ctc_loss
    g = torch.Generator().manual_seed(T + B)
    lp_btv = torch.randn(B, T, V, generator=g).log_softmax(-1)
    tl = torch.randint(max(S // 2, 0), S + 1, (B,), generator=g)
    il = torch.randint(max(T // 2, 2 * S + 1), T + 1, (B,), generator=g)
    tg = torch.full((B, S + 1), 1, dtype=torch.long)
    for b in range(B):
        tg[b, : tl[b]] = torch.randint(3, V, (int(tl[b]),), generator=g)
        tg[b, tl[b]] = 2
    ref_in = lp_btv.transpose(0, 1).clone().requires_grad_(True)
    ref = ref_ctc.ctc_loss_reference(ref_in, il, tg, tl, 0, 1, 2, "sum")
    ref.backward()
    # float64 truth: fp32 log-space CTC (ATen's included) carries alpha/beta of magnitude ~|nll|, so its gradient
    # error grows with T (4.6e-4 at T=250, 2e-3 at T=750 for ATen itself).  The kernel must be as accurate as the
    # reference's own fp32 arithmetic: error vs truth <= 3x ATen-fp32's error vs truth (+5e-5).
    t64 = lp_btv.double().transpose(0, 1).clone().requires_grad_(True)
    ref_ctc.ctc_loss_reference(t64, il, tg, tl, 0, 1, 2, "sum").backward()
    aten_err = (ref_in.grad.double() - t64.grad).abs().max().item()
    d = lp_btv.cuda().requires_grad_(True)
    loss = ctc_loss(d.transpose(0, 1), il.cuda(), tg.cuda(), tl, reduction="sum")
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 2e-5 * abs(ref.item()), (loss.item(), ref.item())
    my_err = (d.grad.transpose(0, 1).cpu().double() - t64.grad).abs().max().item()
    assert my_err <= 3 * aten_err + 5e-5, (my_err, aten_err)
    # size-independent properties: zero gradient beyond each utterance's length; rows sum to ~0 (softmax - occupancy)
    gr = d.grad.cpu()
    for b in range(B):
        assert gr[b, il[b]:].abs().max().item() == 0 if il[b] < T else True
    assert gr.sum(-1).abs().max().item() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("T,B,S,V", [(50, 8, 10, 32), (749, 8, 150, 32), (300, 5, 40, 100), (1500, 16, 300, 32), (40, 3, 0, 8),
                                     (600, 4, 400, 29)])
def test_cuda_ctc_from_logits_matches_log_softmax_then_ctc(T, B, S, V):
    """fused log_softmax + CTC (wav2vec2.py:770 + ctc.py:197): loss and d loss / d logits against float64
    log_softmax -> ctc_loss autograd, through the tagging the acoustic model uses (`a8_logits` on the log-probs)"""
    from audio8_b200.ctc import ctc_loss
    from audio8_b200.functional import LogSoftmaxFn
    g = torch.Generator().manual_seed(3 * T + B)
    logits = torch.randn(B, T, V, generator=g) * 2.0
    tl = torch.randint(max(S // 2, 0), S + 1, (B,), generator=g)
    il = torch.randint(max(T // 2, min(2 * S + 1, T)), T + 1, (B,), generator=g)
    tg = torch.full((B, S + 1), 1, dtype=torch.long)
    for b in range(B):
        tg[b, : tl[b]] = torch.randint(3, V, (int(tl[b]),), generator=g)
    x64 = logits.double().clone().requires_grad_(True)
    ref = ref_ctc.ctc_loss_reference(torch.log_softmax(x64, -1).transpose(0, 1), il, tg, tl, 0, 1, 2, "sum")
    ref.backward()
    x32 = logits.clone().requires_grad_(True)
    ref32 = ref_ctc.ctc_loss_reference(torch.log_softmax(x32, -1).transpose(0, 1), il, tg, tl, 0, 1, 2, "sum")
    ref32.backward()
    aten_err = (x32.grad.double() - x64.grad).abs().max().item()
    d = logits.cuda().requires_grad_(True)
    lp = LogSoftmaxFn.apply(d)
    lp.a8_logits = d  # what Wav2Vec2AcousticModel.forward does
    loss = ctc_loss(lp.transpose(1, 0), il.cuda(), tg.cuda(), tl, reduction="sum")
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 2e-5 * abs(ref.item()) + 1e-4, (loss.item(), ref.item())
    my_err = (d.grad.cpu().double() - x64.grad).abs().max().item()
    assert my_err <= 3 * aten_err + 5e-5, (my_err, aten_err)
    gr = d.grad.cpu()
    for b in range(B):
        assert gr[b, il[b]:].abs().max().item() == 0 if il[b] < T else True
    assert gr.sum(-1).abs().max().item() < 1e-3  # softmax - occupancy sums to zero over classes
    # an untagged / re-laid-out tensor takes the plain log-prob path and agrees
    d2 = logits.cuda().requires_grad_(True)
    loss2 = ctc_loss(torch.log_softmax(d2, -1).transpose(1, 0), il.cuda(), tg.cuda(), tl, reduction="sum")
    loss2.backward()
    assert abs(loss2.item() - loss.item()) <= 2e-5 * abs(loss.item()) + 1e-4
    assert (d2.grad - d.grad).abs().max().item() <= 3 * aten_err + 1e-4


def test_ctc_from_logits_host_logic_cpu(emu_backend):
    """the tagging / view recognition of the fused path, through the ABI emulation"""
    from audio8_b200.ctc import _fused_logits, ctc_loss
    from audio8_b200.functional import LogSoftmaxFn
    g = torch.Generator().manual_seed(0)
    logits = (torch.randn(3, 30, 16, generator=g)).requires_grad_(True)
    lp = LogSoftmaxFn.apply(logits)
    lp.a8_logits = logits
    assert _fused_logits(lp.transpose(1, 0)) is logits
    assert _fused_logits(lp.transpose(1, 0)[1:]) is None and _fused_logits(lp.transpose(1, 0).contiguous()) is None
    assert _fused_logits(lp) is None  # [B,T,V] order is not what the loss takes
    tg = torch.randint(3, 16, (3, 5), generator=g)
    tl = torch.full((3,), 5, dtype=torch.long)
    il = torch.tensor([30, 25, 28])
    loss = ctc_loss(lp.transpose(1, 0), il, tg, tl)
    loss.backward()
    x = logits.detach().clone().requires_grad_(True)
    ref = ref_ctc.ctc_loss_reference(torch.log_softmax(x, -1).transpose(0, 1), il, tg, tl, 0, 1, 2, "sum")
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item())
    assert (logits.grad - x.grad).abs().max().item() < 1e-4


def _greedy_reference(lp_btv, lengths, blank):
    """the reference's own ops (ctc.py:161-162)"""
    out = []
    for lp, n in zip(lp_btv, lengths):
        toks = lp[: int(n)].argmax(dim=-1).unique_consecutive()
        out.append(toks[toks != blank].tolist())
    return out


def _greedy_inputs(B, T, V, seed):
    g = torch.Generator().manual_seed(seed)
    # peaky, run-heavy posteriors (repeats and blanks matter) with exact ties sprinkled in
    base = torch.randint(0, V, (B, (T + 3) // 4), generator=g).repeat_interleave(4, dim=1)[:, :T]
    lp = torch.randn(B, T, V, generator=g) * 0.3
    lp.scatter_add_(2, base[..., None], torch.full((B, T, 1), 2.0))
    lp[:, ::7, 3] = lp[:, ::7].max(-1).values  # ties: the first maximal index must win
    lengths = torch.randint(T // 2, T + 1, (B,), generator=g)
    lengths[0] = T
    return lp.log_softmax(-1), lengths


def test_greedy_decode_host_logic_cpu(emu_backend):
    from audio8_b200.ctc import greedy_decode
    lp, lengths = _greedy_inputs(3, 57, 8, 0)
    assert greedy_decode(lp, lengths, blank=0) == _greedy_reference(lp, lengths, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,V", [(4, 99, 32), (8, 749, 32), (3, 1500, 100), (2, 1, 5), (64, 750, 32)])
def test_cuda_greedy_decode_bit_exact(B, T, V):
    """integer alignments: identical to the reference's argmax / unique_consecutive / blank removal on the same
    log-probs, incl. ties, ragged lengths and the non-contiguous [T,B,V] -> [B,T,V] view the trainer produces"""
    from audio8_b200.ctc import greedy_decode
    lp, lengths = _greedy_inputs(B, T, V, B + T)
    want = _greedy_reference(lp, lengths, 0)
    assert greedy_decode(lp.cuda(), lengths, blank=0) == want
    tbv = lp.transpose(0, 1).contiguous().cuda()  # [T,B,V] storage seen as [B,T,V]
    assert greedy_decode(tbv.transpose(0, 1), lengths.cuda(), blank=0) == want
    assert greedy_decode(lp.cuda(), None, blank=0) == _greedy_reference(lp, [T] * B, 0)

