"""Data-parallel wrapper for the drop-in modules: one process per GPU, batch sharded by utterance, ONE exchange step
per optimizer step (the gradient average) — the same contract as `torch.nn.parallel.DistributedDataParallel`, which the
reference's trainers construct (`pretrain.py:158`, `train.py:266`) and which keeps working with these modules.

What differs from DistributedDataParallel is where the gradients live.  Every backward kernel writes its parameter
gradients straight into a persistent, contiguous fp32 **gradient arena** whose layout is planned once from the module
structure (identical on every rank), so `param.grad` are views of the arena and the all-reduce runs in place: no
per-parameter hook, no bucket copy kernels (DDP launched ~200 of them per step here) and two NCCL calls per step — the
transformer-layer region (89 % of the bytes of wav2vec2-base), started as soon as the encoder's backward has been
enqueued so that it overlaps the conv feature encoder's backward, and the rest (conv stack, projections, quantizer) when
backward ends.

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
    """One persistent fp32 buffer holding every trainable parameter's gradient at a FIXED offset, planned once from the
    module structure (never from the order in which backward happens to reach the layers: with LayerDrop each rank drops
    different layers, and a lazily assigned layout would differ between ranks and corrupt the in-place all-reduce).
    Layout: [transformer layers: one block per layer, last layer first | heads, quantizer, positional conv: everything
    whose backward runs before or inside the encoder's | conv feature encoder, its LayerNorm, input projection, mask
    embedding].  The first two regions are complete when the gradient w.r.t. the encoder's input exists (the modules build
    the encoder BEFORE the quantizer branch when an arena is active, so autograd runs that branch first) and are reduced
    under the conv stack's backward; the last region (5 % of the bytes) is reduced when backward ends."""

    LATE = ("feature_extractor", "layer_norm", "proj_to_input", "mask_emb")  # gradients that complete after the encoder's
    _serial = 0

    def __init__(self, module, device):
        from .wav2vec2 import AudioTransformerEncoder
        plan, covered = [], set()
        for m in module.modules():
            if isinstance(m, AudioTransformerEncoder):
                for layer in reversed(list(m.transformer.encoders)):
                    flat = layer.flat()
                    if not all(p.requires_grad for p in flat):
                        continue
                    plan.append((ops.grad_key(flat[0]), sum(p.numel() for p in flat)))
                    covered.update(id(p) for p in flat)
        named = [(k, p) for k, p in module.named_parameters() if p.requires_grad and id(p) not in covered]
        late = lambda k: any(part in self.LATE for part in k.split("."))
        mid = [(ops.grad_key(p), p.numel()) for k, p in named if not late(k)]
        rest = [(ops.grad_key(p), p.numel()) for k, p in named if late(k)]
        self.slots = {}
        off = 0
        for part in (plan, mid, rest):
            for key, n in part:
                off = (off + 63) & ~63  # 256-byte alignment
                self.slots[key] = (off, n)
                off += n
            if part is mid:
                self.early = off  # end of the region that is complete when the encoder's backward has been enqueued
        self.used = off
        self.buf = torch.zeros(max(off, 1), dtype=torch.float32, device=device)
        GradArena._serial += 1
        self.serial = GradArena._serial  # part of the CUDA-graph keys: a graph captured under this arena writes into it

    def take(self, key, numel, zero=True):
        """the block planned for `key` (same storage on every step and every rank), zeroed unless zero=False; None when
        the key has no block or the size differs (the caller allocates normally and the gradient is reduced separately)"""
        cb = self.__dict__.get("on_first_take")
        if cb is not None:
            self.on_first_take = None
            cb()
        slot = self.slots.get(key)
        if slot is None or slot[1] != numel:
            return None
        out = self.buf[slot[0]:slot[0] + numel]
        if zero:
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
        self._early = None  # work handle of the all-reduce of the transformer-layer region started inside backward
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
            self._arena = GradArena(self.module, self._params[0].device)
        if self._arena is not None:
            self._arena.on_first_take = None  # armed after forward: a CUDA-graph capture inside forward must not trigger it
        ops.set_grad_arena(self._arena if fresh else None, self._encoder_done if (fresh and self.overlap) else None)
        try:
            out = self.module(*args, **kwargs)
        finally:
            ops.set_grad_arena(None, None, keep_for_backward=True)
        if sync:
            self._callback_queued = False
            if fresh and self._arena is not None:
                self._arena.on_first_take = self._queue_finish  # the first gradient block handed out in backward
            for t in (out if isinstance(out, (tuple, list)) else (out,)):
                # the loss may bypass an output and differentiate a tensor it carries instead (the CTC loss works from the
                # logits tagged onto the log-probs: wav2vec2.Wav2Vec2AcousticModel.forward): hook both
                for u in (t, getattr(t, "a8_logits", None)):
                    if isinstance(u, torch.Tensor) and u.requires_grad:
                        u.register_hook(self._queue_callback)
        return out

    # ------------------------------------------------------------------------------------------------ backward side
    def _queue_callback(self, grad):
        self._queue_finish()
        return grad

    def _queue_finish(self):
        """(inside backward) make sure `_finish` runs when this backward pass ends"""
        if not self._callback_queued:
            self._callback_queued = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)

    def _all_reduce(self, t, async_op):
        if dist.get_backend(self.pg) == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.pg, async_op=async_op)
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.pg, async_op=False)  # gloo (CPU tests): no AVG
        t.div_(self.world)
        return w if async_op else None

    def _encoder_done(self):
        """called (on the autograd thread) when the gradient w.r.t. the encoder's input exists, i.e. every transformer
        layer has put its gradients into the arena: start reducing that region under the rest of backward"""
        a = self._arena
        if a is None or a.early == 0 or self._early is not None:
            return
        self._early = self._all_reduce(a.buf[:a.early], async_op=True) or True

    def _finish(self):
        a = self._arena
        arena_on = ops.grad_arena_for_backward() is a and a is not None
        works = []
        if arena_on:
            lo = a.early if self._early is not None else 0  # whatever has not been started yet: one call, in place
            if a.used > lo:
                works.append(self._all_reduce(a.buf[lo:a.used], async_op=True))
            ptr0, ptr1 = a.buf.data_ptr(), a.buf.data_ptr() + a.buf.numel() * 4
        rest = []
        for p in self._params:
            g = p.grad
            if g is None or (arena_on and ptr0 <= g.data_ptr() < ptr1):
                continue  # no gradient, or one that lives in the arena (reduced in place)
            rest.append(g)
        if rest:  # gradients produced outside the arena (gradient accumulation, or a parameter the plan does not know)
            flat = torch.cat([g.reshape(-1) for g in rest])
            self._all_reduce(flat, async_op=False)
            torch._foreach_copy_(rest, [f.view_as(g) for f, g in zip(flat.split([g.numel() for g in rest]), rest)])
        if self._early is not None and self._early is not True:
            works.append(self._early)
        for w in works:
            if w is not None:
                w.wait()  # the compute stream waits for the collective; the host does not
        self._early = None
        ops.set_grad_arena(None, None)
