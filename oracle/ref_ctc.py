"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's CTC loss path.  Never imported by the product.

* `ctc_loss_reference`  follows `/root/reference/audio8/ctc.py:186-206` literally: strip PAD/EOS from the padded
  targets, then call the installed `torch.nn.functional.ctc_loss` (the exact function the reference calls) —
  this is the pin for the CUDA kernel.
* `ctc_numpy`  is an independent float64 restatement of the alpha/beta recursion and of PyTorch's gradient
  convention `exp(lp) - occupancy`; `tests/test_oracle.py` checks it against
  `F.ctc_loss` so the closed form the kernel implements is itself pinned.
* `greedy_decode`  follows `ctc.py:161-162` (argmax, unique_consecutive, drop blank).
"""
import numpy as np
import torch
import torch.nn.functional as F


def strip_targets(targets, pad, eos):
    """ctc.py:193-194: row-major concatenation of the entries that are neither PAD nor EOS."""
    keep = (targets != pad) & (targets != eos)
    return targets.masked_select(keep)


def ctc_loss_reference(log_prob, input_lengths, targets, target_lengths, blank=0, pad=1, eos=2,
                       reduction="sum", zero_infinity=True):
    flat = strip_targets(targets, pad, eos)
    return F.ctc_loss(log_prob, flat, input_lengths, target_lengths, blank=blank, reduction=reduction,
                      zero_infinity=zero_infinity)


def _lse(*xs):
    m = max(xs)
    if m == -np.inf:
        return -np.inf
    return m + np.log(sum(np.exp(x - m) for x in xs))


def ctc_numpy(lp, input_lengths, labels, blank=0):
    """lp [T,B,V] float64 log-probs; labels: list of 1-D int arrays.  Returns (nll [B], grad [T,B,V]) with
    grad = exp(lp) - occupancy for t < Tb and 0 for t >= Tb; infeasible rows give nll=inf, grad 0."""
    T, B, V = lp.shape
    nll = np.zeros(B)
    grad = np.zeros_like(lp)
    for b in range(B):
        Tb = int(input_lengths[b])
        l = np.asarray(labels[b], dtype=np.int64)
        S = len(l)
        E = 2 * S + 1
        ext = np.full(E, blank, dtype=np.int64)
        ext[1::2] = l
        skip = np.zeros(E, dtype=bool)
        for s in range(2, E):
            skip[s] = ext[s] != blank and ext[s] != ext[s - 2]
        a = np.full((Tb, E), -np.inf)
        a[0, 0] = lp[0, b, blank]
        if E > 1:
            a[0, 1] = lp[0, b, ext[1]]
        for t in range(1, Tb):
            for s in range(E):
                terms = [a[t - 1, s]]
                if s >= 1:
                    terms.append(a[t - 1, s - 1])
                if skip[s]:
                    terms.append(a[t - 1, s - 2])
                a[t, s] = lp[t, b, ext[s]] + _lse(*terms)
        ll = _lse(a[Tb - 1, E - 1], a[Tb - 1, E - 2]) if E > 1 else a[Tb - 1, 0]
        nll[b] = -ll
        if not np.isfinite(ll):
            continue
        bt = np.full((Tb, E), -np.inf)
        bt[Tb - 1, E - 1] = lp[Tb - 1, b, blank]
        if E > 1:
            bt[Tb - 1, E - 2] = lp[Tb - 1, b, ext[E - 2]]
        for t in range(Tb - 2, -1, -1):
            for s in range(E):
                terms = [bt[t + 1, s]]
                if s + 1 < E:
                    terms.append(bt[t + 1, s + 1])
                if s + 2 < E and skip[s + 2]:
                    terms.append(bt[t + 1, s + 2])
                bt[t, s] = lp[t, b, ext[s]] + _lse(*terms)
        for t in range(Tb):
            occ = np.zeros(V)
            for s in range(E):
                v = a[t, s] + bt[t, s]
                if v > -np.inf:
                    occ[ext[s]] += np.exp(v - lp[t, b, ext[s]] - ll)
            grad[t, b] = np.exp(lp[t, b]) - occ
    return nll, grad


def greedy_decode(log_probs, input_lengths, blank=0):
    """ctc.py:156-162: per utterance argmax over classes, collapse repeats, drop blanks."""
    out = []
    for lp, n in zip(log_probs, input_lengths):
        toks = lp[: int(n)].argmax(-1).unique_consecutive()
        out.append(toks[toks != blank].tolist())
    return out
