"""Drop-in for the CTC loss of `audio8/ctc.py` (reference: /root/reference/audio8/ctc.py:186-206).

`CTCLoss` keeps the reference's constructor and call signature; the arithmetic runs in hand-written
sm_100a kernels (audio8_b200/csrc/ctc.cu) instead of ATen's ctc_loss kernels.  Blank / PAD / EOS ids are
read from `Offsets` at call time, like the reference does (`ctc.py:193,202`; `train.py:22-27` re-points them).
"""
import torch

from . import ops
from .functional import _take_saved


class Offsets:
    """Same defaults as eight_mile.utils.Offsets; train.py sets GO=0, PAD=1 at import time."""

    PAD, GO, EOS, UNK, OFFSET = 0, 1, 2, 3, 4
    VALUES = ["<PAD>", "<GO>", "<EOS>", "<UNK>"]


def _offsets():
    try:  # when running inside the reference's trainers, honour their (mutated) Offsets
        from eight_mile.utils import Offsets as O
        return O
    except Exception:
        return Offsets


class _CTCFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_prob, input_lengths, targets, target_lengths, blank, pad, eos, mean, zero_infinity):
        be = ops.backend()
        dev = log_prob.device
        lp = log_prob if log_prob.dtype == torch.float32 else log_prob.float()
        B, S = targets.shape
        # max target length: free if the lengths live on the host (as in train.py), else bounded by S (no sync)
        max_S = int(target_lengths.max()) if not target_lengths.is_cuda else S
        max_S = min(max_S, S) if S > 0 else 0
        tl = target_lengths.to(device=dev, dtype=torch.int64, non_blocking=True)
        il = torch.as_tensor(input_lengths).to(device=dev, dtype=torch.int64, non_blocking=True)
        tg = targets.to(device=dev, dtype=torch.int64, non_blocking=True)
        flat, off, tl32, il32 = be.ctc_prep(tg, pad, eos, tl, il)
        loss, nll, alpha, beta = be.ctc_forward(lp, flat, off, tl32, il32, max_S, blank, mean, zero_infinity)
        ctx.saved = (lp, flat, off, tl32, il32, alpha, beta, nll)
        ctx.cfg = (max_S, blank, mean, zero_infinity, log_prob.dtype)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lp, flat, off, tl32, il32, alpha, beta, nll = _take_saved(ctx)
        max_S, blank, mean, zero_infinity, dtype = ctx.cfg
        grad = ops.backend().ctc_backward(lp, flat, off, tl32, il32, max_S, blank, alpha, beta, nll, grad_out, mean,
                                          zero_infinity)
        return grad.to(dtype), None, None, None, None, None, None, None, None


def ctc_loss(log_prob, input_lengths, targets, target_lengths, blank=0, pad=1, eos=2, reduction="sum",
             zero_infinity=True):
    """log_prob [T,B,V] (any strides, e.g. the transposed view train.py:39 passes), targets [B,S] padded."""
    if reduction not in ("sum", "mean"):
        raise ValueError(f"reduction {reduction!r} not supported (the reference uses 'sum', ctc.py:187)")
    return _CTCFunction.apply(log_prob, input_lengths, targets, target_lengths, int(blank), int(pad), int(eos),
                              reduction == "mean", bool(zero_infinity))


class CTCLoss(torch.nn.Module):
    """Same interface as audio8.ctc.CTCLoss (ctc.py:186-206)."""

    def __init__(self, zero_infinity=True, reduction_type="sum"):
        super().__init__()
        self.zero_infinity = zero_infinity
        self.reduction_type = reduction_type

    def forward(self, log_prob, input_lengths, targets, target_lengths):
        O = _offsets()
        return ctc_loss(log_prob, input_lengths, targets, target_lengths, blank=O.GO, pad=O.PAD, eos=O.EOS,
                        reduction=self.reduction_type, zero_infinity=self.zero_infinity)


def greedy_decode(log_probs, input_lengths=None, blank=None):
    """Best-path decode on the device: per utterance `log_probs[b, :len].argmax(-1).unique_consecutive()` with the blank
    removed — exactly the token sequence `ctc_metrics` feeds to editdistance (reference ctc.py:161-162), without moving the
    [B,T,V] log-probs to the host (`train.py:46`).  log_probs [B,T,V]; returns a list of B python lists of ints."""
    if blank is None:
        blank = _offsets().GO
    lp = log_probs.detach()
    if lp.dtype != torch.float32:
        lp = lp.float()
    il = None
    if input_lengths is not None:
        il = torch.as_tensor(input_lengths).to(device=lp.device, dtype=torch.int32)
    ids, lens = ops.backend().ctc_greedy(lp, il, int(blank))
    ids, lens = ids.cpu(), lens.cpu().tolist()
    return [ids[b, : lens[b]].tolist() for b in range(lp.shape[0])]

