"""Optimizer side of the training step: gradient-norm clipping + AdamW (+ gradient unscale) as two
multi-tensor launches over every parameter, with `torch.optim.AdamW`'s interface and arithmetic.

Reference call sites (under /root/reference/audio8): `pretrain.py:182-184`
    torch.nn.utils.clip_grad_norm_(model.parameters(), args.clip); optimizer.step(); optimizer.zero_grad()
and `train.py:323-325` (the same preceded by `optimizer.scale_grads(num_gpus / batch_size)`), where `optimizer` is
eight_mile's `OptimizerManager` around `torch.optim.AdamW`.  Drop-in use:

    opt = FusedAdamW(model.parameters(), lr=..., weight_decay=...)      # instead of torch.optim.AdamW
    opt.step(clip=args.clip, grad_scale=num_gpus / batch_size)          # clip + unscale + update: 2 launches
or, with the trainers untouched, `clip_grad_norm_` below instead of torch's and a plain `opt.step()`.

The kernels (csrc/optim.cu) read each gradient once for the norm (4 B / element) and once for the update, which reads
and writes p, exp_avg, exp_avg_sq (28 B / element; HBM-bound: 95 M parameters = 3.0 GB per step).  The squared norm is
reduced deterministically (per-chunk partials folded in double precision by every CTA of the update kernel): no
atomics, no host synchronisation, nothing to zero.
"""
import math

import numpy as np
import torch

from . import ops

CHUNK = 65536  # elements per CTA pass (multiple of 4)


class _Plan:
    """chunk -> (tensor, offset) map of a parameter list: depends only on the numels, built once"""

    def __init__(self, params, device):
        self.params = params
        ct, co = [], []
        for i, p in enumerate(params):
            for off in range(0, p.numel(), CHUNK):
                ct.append(i)
                co.append(off)
        self.n_chunks = len(ct)
        self.chunk_tensor = torch.tensor(ct, dtype=torch.int32).to(device)
        self.chunk_off = torch.tensor(co, dtype=torch.int64).to(device)
        self.partials = torch.empty(max(self.n_chunks, 1), dtype=torch.float32, device=device)
        self.table_np = np.zeros((len(params), 6), dtype=np.int64)
        self.table = torch.empty((len(params), 6), dtype=torch.int64, device=device)
        self.sig = None

    def refresh(self, state_of, copies):
        """rewrite the pointer table when any address changed (gradients are fresh allocations on most steps; under the
        data-parallel wrapper they are fixed views of its arena and the upload is skipped)"""
        t = self.table_np
        for i, p in enumerate(self.params):
            g = p.grad
            st = state_of(p) if state_of is not None else None
            t[i, 0] = p.data_ptr()
            t[i, 1] = g.data_ptr() if g is not None else 0
            t[i, 2] = st["exp_avg"].data_ptr() if st is not None else 0
            t[i, 3] = st["exp_avg_sq"].data_ptr() if st is not None else 0
            t[i, 4] = p.numel()
            c = copies.get(p) if copies else None
            t[i, 5] = c.data_ptr() if c is not None else 0
        sig = t.tobytes()
        if sig != self.sig:
            self.sig = sig
            if self.table.is_cuda:
                from .wav2vec2 import _to_device
                self.table = _to_device(t, self.table.device)
            else:
                self.table = torch.from_numpy(t.copy())
        return self.table


def _check(params):
    for p in params:
        g = p.grad
        if not (p.dtype == torch.float32 and p.is_contiguous()):
            raise ValueError("FusedAdamW / clip_grad_norm_: parameters must be contiguous fp32 tensors")
        if g is not None and not (g.dtype == torch.float32 and g.is_contiguous()):
            raise ValueError("FusedAdamW / clip_grad_norm_: gradients must be contiguous fp32 (dense) tensors")


_CLIP_PLANS = {}


def clip_grad_norm_(parameters, max_norm, norm_type=2.0):
    """Drop-in for `torch.nn.utils.clip_grad_norm_` (pretrain.py:182, train.py:324): scales the gradients in place so that
    their global 2-norm is at most `max_norm`, returns the norm before clipping (a device scalar: no host sync)."""
    if float(norm_type) != 2.0:
        raise NotImplementedError("the reference clips the 2-norm (torch's default)")
    params = [p for p in (parameters if not isinstance(parameters, torch.Tensor) else [parameters]) if p.requires_grad]
    if not params:
        return torch.zeros(())
    _check(params)
    key = tuple(id(p) for p in params)
    plan = _CLIP_PLANS.get(key)
    if plan is None:
        _CLIP_PLANS.clear()
        plan = _CLIP_PLANS[key] = _Plan(params, params[0].device)
    be = ops.backend()
    table = plan.refresh(None, None)
    total = torch.empty((), dtype=torch.float32, device=params[0].device)
    be.optim_grad_sqnorm(table, plan.chunk_tensor, plan.chunk_off, CHUNK, plan.partials)
    be.optim_adamw(table, plan.chunk_tensor, plan.chunk_off, CHUNK, plan.partials, float(max_norm), 1.0, 0.0, 0.0, 0.0,
                   0.0, 0.0, 1.0, 1.0, True, total)
    return total


class FusedAdamW(torch.optim.Optimizer):
    """`torch.optim.AdamW` (amsgrad=False, maximize=False) with the update of ALL parameters in one launch, optionally
    fused with gradient-norm clipping (`step(clip=...)`) and gradient unscaling (`step(grad_scale=...)`; `scale_grads(s)`,
    eight_mile OptimizerManager's name for it, scales in place with one launch).  State keys match torch's
    (`step`, `exp_avg`, `exp_avg_sq`), so `state_dict()` round-trips with `torch.optim.AdamW`."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._plans = {}
        self.operand_copies = {}  # parameter -> persistent bf16 buffer refreshed by the update kernel (optional)
        self.last_grad_norm = None

    def scale_grads(self, s):
        """eight_mile OptimizerManager.scale_grads (train.py:323): every gradient *= s, in place, one launch (so that a
        later `clip_grad_norm_` sees the scaled gradients exactly as in the reference's call order).  `step(grad_scale=s)`
        is the fused alternative that never writes the gradients back."""
        params = [p for g in self.param_groups for p in g["params"] if p.requires_grad]
        if not params:
            return
        _check(params)
        plan = self._plan(("all",), params)
        table = plan.refresh(None, None)
        ops.backend().optim_adamw(table, plan.chunk_tensor, plan.chunk_off, CHUNK, None, 0.0, float(s), 0.0, 0.0, 0.0, 0.0,
                                  0.0, 1.0, 1.0, True, None)

    def register_operand_copy(self, param, bf16_buffer):
        """the update kernel also writes bf16(param) into `bf16_buffer` (same numel): the GEMM operand copy a forward pass
        would otherwise re-cast"""
        assert bf16_buffer.dtype == torch.bfloat16 and bf16_buffer.numel() == param.numel() and bf16_buffer.is_contiguous()
        self.operand_copies[param] = bf16_buffer
        for plan in self._plans.values():
            plan.sig = None

    @torch.no_grad()
    def step(self, closure=None, clip=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        be = ops.backend()
        gscale = float(grad_scale)
        groups = [([p for p in g["params"] if p.requires_grad], g) for g in self.param_groups]
        partial_sets = []
        if clip is not None and clip > 0:  # the norm runs over ALL groups' gradients, like clip_grad_norm_(model.parameters())
            allp = [p for ps, _ in groups for p in ps]
            plan = self._plan(("all",), allp)
            _check(allp)
            table = plan.refresh(None, None)
            be.optim_grad_sqnorm(table, plan.chunk_tensor, plan.chunk_off, CHUNK, plan.partials)
            partial_sets = plan.partials
            self.last_grad_norm = torch.empty((), dtype=torch.float32, device=allp[0].device)
        for gi, (ps, g) in enumerate(groups):
            if not ps:
                continue
            _check(ps)
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            # one step counter per group (torch keeps one per parameter; they only differ for parameters that skip steps)
            stepped = [p for p in ps if p.grad is not None]
            if not stepped:
                continue
            t = None
            for p in stepped:
                self.state[p]["step"] += 1
                t = float(self.state[p]["step"]) if t is None else t
            if any(float(self.state[p]["step"]) != t for p in stepped):
                raise RuntimeError("FusedAdamW: parameters of one group are at different step counts (a parameter received "
                                   "gradients on some steps only); put such parameters into their own param group")
            b1, b2 = g["betas"]
            bc1 = 1.0 - b1 ** t
            bc2s = math.sqrt(1.0 - b2 ** t)
            plan = self._plan(("g", gi), ps)
            table = plan.refresh(lambda p: self.state[p], self.operand_copies)
            has_clip = clip is not None and clip > 0
            be.optim_adamw(table, plan.chunk_tensor, plan.chunk_off, CHUNK, partial_sets if has_clip else None,
                           float(clip) if has_clip else 0.0, gscale, float(g["lr"]), float(b1), float(b2), float(g["eps"]),
                           float(g["weight_decay"]), bc1, bc2s, False, self.last_grad_norm if has_clip else None)
            for p in stepped:
                # the kernel wrote through raw pointers: tell autograd (and the operand caches keyed on the version
                # counter, functional.operands) that the parameter changed
                torch.autograd.graph.increment_version(p)
        return loss

    def _plan(self, key, params):
        key = key + tuple(id(p) for p in params)
        plan = self._plans.get(key)
        if plan is None:
            plan = self._plans[key] = _Plan(params, params[0].device)
        return plan
