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
import os
import warnings

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib, ops


class SwitchAllReduce:
    """In-place averaging all-reduce of ranges of ONE fp32 buffer through the NVSwitch's multicast + in-fabric reduction
    (csrc/allreduce_mc.cu).  The buffer is a symmetric-memory allocation that every rank of the group maps into one
    multicast window (torch.distributed._symmetric_memory does the handle exchange: plumbing); the reduction itself is
    this repo's kernel, bracketed by the symmetric-memory barrier, on a side stream so that it runs under backward.
    Construction is a collective; it raises when the fabric offers no multicast (no NVSwitch, or NVLS disabled)."""

    def __init__(self, numel, device, group):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.buf = symm.empty(numel, dtype=torch.float32, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.mc = int(self.hdl.multicast_ptr)
        if not self.mc:
            raise RuntimeError("symmetric memory without a multicast mapping")
        self.buf.zero_()
        self.stream = torch.cuda.Stream(device)
        # CTAs that saturate the links (scripts/switch_probe.py: 24 at world 2, 8 at world 8 - every ld_reduce fans out to
        # world GPUs inside the switch); fewer CTAs leave more SMs to the backward kernels running next to it
        self.ctas = int(os.environ.get("A8_ALLREDUCE_CTAS", "0")) or (24 if self.world <= 2 else 16 if self.world <= 4 else 8)
        self._lib = _lib.load()

    def start(self, lo, hi):
        """queue the all-reduce of elements [lo, hi) (multiples of 4) behind everything already on the current stream;
        returns an object whose wait() makes the current stream wait for the result (the host never blocks)"""
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            self.hdl.barrier(channel=0)  # every rank's gradients of this range are written
            _lib.check(self._lib.a8_allreduce_mc(self.mc, lo, hi, self.rank, self.world, 1.0 / self.world, self.ctas,
                                                 self.stream.cuda_stream), "a8_allreduce_mc")
            self.hdl.barrier(channel=1)  # every rank's slice has landed everywhere
        return self

    def wait(self):
        torch.cuda.current_stream().wait_stream(self.stream)


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

    def __init__(self, module, device, group=None):
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
        size = (max(off, 1) + 63) & ~63
        self.switch = self._switch(size, device, group)
        self.buf = self.switch.buf if self.switch is not None else torch.zeros(size, dtype=torch.float32, device=device)
        GradArena._serial += 1
        self.serial = GradArena._serial  # part of the CUDA-graph keys: a graph captured under this arena writes into it

    @staticmethod
    def _switch(size, device, group):
        """the NVSwitch all-reduce for this arena when every rank can have it (NCCL group on CUDA devices with a
        multicast-capable fabric, world size >= 4; A8_ALLREDUCE=nccl / switch force one or the other), else None: NCCL's
        all-reduce on a plain buffer"""
        if not (dist.is_initialized() and torch.device(device).type == "cuda" and dist.get_backend(group) == "nccl"
                and dist.get_world_size(group) > 1):
            return None
        # measured (profiles/r02_dp.md): at world 2 a ring moves as many bytes as the switch path and NCCL's kernel is
        # the faster one (698 vs 902 us for 362 MB); from world 4 on the in-switch reduction wins (776 vs 1022 us at 8)
        mode = os.environ.get("A8_ALLREDUCE", "auto")
        if mode == "nccl" or (mode == "auto" and dist.get_world_size(group) < 4):
            return None
        sw, err = None, None
        try:
            sw = SwitchAllReduce(size, device, group)
        except Exception as e:  # noqa: BLE001 - any failure means "not available here"; the ranks agree below
            err = e
        ok = torch.tensor([1 if sw is not None else 0], device=device, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 1:
            return sw
        if dist.get_rank(group) == 0:
            warnings.warn(f"audio8_b200: NVSwitch multicast all-reduce unavailable ({err!r}); using NCCL")
        return None

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
    def __init__(self, module, process_group=None, overlap=True, force_sync=False):
        super().__init__()
        self.force_sync = force_sync  # diagnostics: run the arena + exchange machinery even with one rank
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
        sync = self.require_sync and (self.world > 1 or self.force_sync) and torch.is_grad_enabled()
        fresh = sync and all(p.grad is None for p in self._params)
        if fresh and self._arena is None and self._params:
            self._arena = GradArena(self.module, self._params[0].device, self.pg)
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

    def _reduce_range(self, a, lo, hi):
        """start the in-place average of arena elements [lo, hi); returns something with wait() (or None)"""
        if a.switch is not None:
            return a.switch.start(lo, (hi + 3) & ~3)
        return self._all_reduce(a.buf[lo:hi], async_op=True)

    def _encoder_done(self):
        """called (on the autograd thread) when the gradient w.r.t. the encoder's input exists, i.e. every transformer
        layer has put its gradients into the arena: start reducing that region under the rest of backward"""
        a = self._arena
        if a is None or a.early == 0 or self._early is not None:
            return
        self._early = self._reduce_range(a, 0, a.early) or True

    def _finish(self):
        a = self._arena
        arena_on = ops.grad_arena_for_backward() is a and a is not None
        works = []
        if arena_on:
            lo = a.early if self._early is not None else 0  # whatever has not been started yet: one call, in place
            if a.used > lo:
                works.append(self._reduce_range(a, lo, a.used))
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
