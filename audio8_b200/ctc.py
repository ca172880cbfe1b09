"""Drop-in for the CTC loss of `audio8/ctc.py` (reference: /root/reference/audio8/ctc.py:186-206).

`CTCLoss` keeps the reference's constructor and call signature; the arithmetic runs in hand-written
sm_100a kernels (audio8_b200/csrc/ctc_loss.cu, ctc.cu) instead of ATen's ctc_loss kernels.  Blank / PAD / EOS ids are
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
    """x: log-probs [T,B,V] (any strides) or, with from_logits, the [T,B,V] view of the classifier's logits [B,T,V]:
    the kernels then normalise the rows themselves and backward returns d loss / d logits directly (the log_softmax
    at wav2vec2.py:770 and its backward never run as separate passes)."""

    @staticmethod
    def forward(ctx, log_prob, input_lengths, targets, target_lengths, blank, pad, eos, mean, zero_infinity,
                from_logits=False):
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
        loss, nll, alpha = be.ctc_forward(lp, flat, off, tl32, il32, max_S, blank, mean, zero_infinity, from_logits)
        ctx.saved = (lp, flat, off, tl32, il32, alpha, nll)
        # the gradient is written in the memory order of the input: [B,T,V]-major for the transposed views train.py passes
        batch_major = lp.dim() == 3 and lp.stride(1) > lp.stride(0)
        ctx.cfg = (max_S, blank, mean, zero_infinity, log_prob.dtype, from_logits, batch_major)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        lp, flat, off, tl32, il32, alpha, nll = _take_saved(ctx)
        max_S, blank, mean, zero_infinity, dtype, from_logits, batch_major = ctx.cfg
        grad = ops.backend().ctc_backward(lp, flat, off, tl32, il32, max_S, blank, alpha, nll, grad_out, mean,
                                          zero_infinity, from_logits, batch_major)
        return grad.to(dtype), None, None, None, None, None, None, None, None, None


def ctc_loss(log_prob, input_lengths, targets, target_lengths, blank=0, pad=1, eos=2, reduction="sum",
             zero_infinity=True):
    """log_prob [T,B,V] (any strides, e.g. the transposed view train.py:39 passes), targets [B,S] padded."""
    if reduction not in ("sum", "mean"):
        raise ValueError(f"reduction {reduction!r} not supported (the reference uses 'sum', ctc.py:187)")
    logits = _fused_logits(log_prob)
    if logits is not None:
        return _CTCFunction.apply(logits.transpose(0, 1), input_lengths, targets, target_lengths, int(blank), int(pad),
                                  int(eos), reduction == "mean", bool(zero_infinity), True)
    return _CTCFunction.apply(log_prob, input_lengths, targets, target_lengths, int(blank), int(pad), int(eos),
                              reduction == "mean", bool(zero_infinity))


def _fused_logits(log_prob):
    """`Wav2Vec2AcousticModel.forward` tags the log-probs it returns with the logits they were computed from
    (`a8_logits`).  When the tensor handed to the loss is that tensor, or the `[T,B,V]` transposed view of it that
    `train.py:39` builds, the loss is computed from the logits and its gradient flows into them directly; anything
    else (sliced, copied, cast, produced elsewhere) takes the plain log-prob path."""
    base = log_prob._base if log_prob._base is not None else log_prob
    logits = getattr(base, "a8_logits", None)
    if logits is None or logits.dtype != torch.float32 or logits.dim() != 3:
        return None
    B, T, V = logits.shape
    same_storage = log_prob.untyped_storage().data_ptr() == base.untyped_storage().data_ptr()
    if not (same_storage and log_prob.storage_offset() == base.storage_offset() and base.is_contiguous()):
        return None
    if tuple(log_prob.shape) == (T, B, V) and log_prob.stride() == (V, T * V, 1):
        return logits
    return None


class CTCLoss(torch.nn.Module):
    """Same interface as audio8.ctc.CTCLoss (ctc.py:186-206)."""

    def __init__(self, zero_infinity=True, reduction_type="sum"):
        super().__init__()
        self.zero_infinity = zero_infinity
        self.reduction_type = reduction_type

    def forward(self, log_prob, input_lengths, targets, target_lengths):
        O = _offsets()
        loss = ctc_loss(log_prob, input_lengths, targets, target_lengths, blank=O.GO, pad=O.PAD, eos=O.EOS,
                        reduction=self.reduction_type, zero_infinity=self.zero_infinity)
        from . import feed
        if feed._DEFERRED:  # the step's forward is enqueued: the input feed may post its next H2D copy now
            feed.flush_deferred()
        return loss


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

