"""Data-parallel wrapper for the drop-in modules: one process per GPU, batch sharded by utterance, ONE exchange step
per optimizer step (the gradient average) — the same contract as `torch.nn.parallel.DistributedDataParallel`, which the
reference's trainers construct (`pretrain.py:158`, `train.py:266`) and which keeps working with these modules.

What differs from DistributedDataParallel is where the gradients live.  The transformer layers (89 % of the parameters
of wav2vec2-base) write their weight gradients straight into a persistent, contiguous fp32 **gradient arena**
(`functional.EncoderFn.backward` takes each layer's accumulator block from it), so `param.grad` of those parameters are
views of the arena and the all-reduce runs in place over one buffer: no per-parameter hook, no bucket copy kernels (DDP
launched ~200 of them per step here) and a single NCCL call, started as soon as the encoder's backward has been
enqueued so that it overlaps the conv feature encoder's backward.  The remaining gradients (conv stack, projections,
quantizer: ~10 % of the bytes) are flattened, reduced and copied back with three launches at the end of backward.

Interface kept: `.module`, `forward`, `no_sync()`, `state_dict()` with the `module.` prefix.  Gradient accumulation
(`no_sync()` micro-steps, or gradients not reset to None) falls back to freshly allocated gradients and the flattening
path for everything, because arena-backed gradients alias the next backward's output.
"""
import contextlib

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops


class GradArena:
    """bump allocator over one persistent fp32 buffer; blocks are keyed (a layer's first parameter) and keep their place"""

    def __init__(self, numel, device):
        self.buf = torch.zeros(numel, dtype=torch.float32, device=device)
        self.slots = {}
        self.used = 0

    def take(self, key, numel):
        """zeroed block of `numel` floats for `key` (same storage on every step), or None when the arena is full"""
        slot = self.slots.get(key)
        if slot is None:
            start = (self.used + 63) & ~63  # 256-byte alignment
            if start + numel > self.buf.numel():
                return None
            slot = self.slots[key] = (start, numel)
            self.used = start + numel
        start, n = slot
        if n != numel:
            return None
        out = self.buf[start:start + n]
        out.zero_()
        return out


class DataParallel(nn.Module):
    def __init__(self, module, process_group=None, overlap=True):
        super().__init__()
        self.module = module
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.require_sync = True
        self.overlap = overlap
        self._params = [p for p in module.parameters() if p.requires_grad]
        self._arena = None
        self._early = None  # (work handle, numel) of the arena all-reduce started inside backward
        self._callback_queued = False
        if self.world > 1:  # every rank starts from rank 0's parameters, like DistributedDataParallel
            with torch.no_grad():
                flat = torch.cat([p.detach().reshape(-1) for p in module.parameters()])
                dist.broadcast(flat, 0, group=process_group)
                o = 0
                for p in module.parameters():
                    p.copy_(flat[o:o + p.numel()].view_as(p))
                    o += p.numel()

    # ------------------------------------------------------------------------------------------------ interface
    @contextlib.contextmanager
    def no_sync(self):
        old, self.require_sync = self.require_sync, False
        try:
            yield
        finally:
            self.require_sync = old

    def forward(self, *args, **kwargs):
        sync = self.require_sync and self.world > 1 and torch.is_grad_enabled()
        fresh = sync and all(p.grad is None for p in self._params)
        if fresh and self._arena is None and self._params:
            n = sum(p.numel() for p in self._params) + 64 * 64
            self._arena = GradArena(n, self._params[0].device)
        ops.set_grad_arena(self._arena if fresh else None, self._encoder_done if (fresh and self.overlap) else None)
        try:
            out = self.module(*args, **kwargs)
        finally:
            ops.set_grad_arena(None, None, keep_for_backward=True)
        if sync:
            self._callback_queued = False
            for t in (out if isinstance(out, (tuple, list)) else (out,)):
                if isinstance(t, torch.Tensor) and t.requires_grad:
                    t.register_hook(self._queue_callback)
        return out

    # ------------------------------------------------------------------------------------------------ backward side
    def _queue_callback(self, grad):
        if not self._callback_queued:
            self._callback_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)
        return grad

    def _all_reduce(self, t, async_op):
        if dist.get_backend(self.pg) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.pg, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg, async_op=False)  # gloo (CPU tests): no AVG
        t.div_(self.world)
        return w if async_op else None

    def _encoder_done(self):
        """called (on the autograd thread) when the gradient w.r.t. the encoder's input exists, i.e. every transformer
        layer has put its gradients into the arena: start reducing them under the rest of backward"""
        a = self._arena
        if a is None or a.used == 0 or self._early is not None:
            return
        self._early = (self._all_reduce(a.buf[:a.used], async_op=True), a.used)

    def _finish(self):
        a = self._arena
        arena_ptr0 = a.buf.data_ptr() if a is not None else 0
        arena_ptr1 = arena_ptr0 + (a.buf.numel() * 4 if a is not None else 0)
        arena_on = ops.grad_arena_for_backward() is a and a is not None
        if arena_on and a.used and self._early is None:
            self._early = (self._all_reduce(a.buf[:a.used], async_op=True), a.used)
        rest = []
        for p in self._params:
            g = p.grad
            if g is None:
                continue
            if arena_on and arena_ptr0 <= g.data_ptr() < arena_ptr1:
                continue  # lives in the arena: reduced in place
            rest.append(g)
        if rest:
            flat = torch.cat([g.reshape(-1) for g in rest])
            self._all_reduce(flat, async_op=False)
            torch._foreach_copy_(rest, [f.view_as(g) for f, g in zip(flat.split([g.numel() for g in rest]), rest)])
        if self._early is not None:
            w = self._early[0]
            if w is not None:
                w.wait()  # the compute stream waits for the collective; the host does not
            self._early = None
        ops.set_grad_arena(None, None)
