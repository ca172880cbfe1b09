"""Shape-keyed CUDA-graph execution of the static-shape segments of the training step.

The step launches ~400 small kernels; enqueueing them from Python costs more host time than the GPU needs to run
them.  The two segments whose shapes depend only on (batch, samples) — the conv feature encoder with its LayerNorm /
projection, and the transformer encoder — are therefore captured once per shape signature (forward and backward
separately, `torch.cuda.make_graphed_callables`) and replayed afterwards; everything data-dependent (span masks,
masked-row gathers, quantizer, contrastive loss) stays eager.  Dropout stays random across replays because kernels
read their per-step seed from device memory (ops.set_seed_source) that torch's graph-safe CUDA generator refreshes.

Policy: a signature is captured when it is seen for the `A8_GRAPH_AFTER`-th time (default 2), at most
`A8_GRAPH_SHAPES` signatures per segment are kept (default 2; further shapes run eagerly), `A8_CUDA_GRAPHS=0`
disables the mechanism.  Captured activations live in the graph's private pool (~4 GB at B=6 x 15 s).
"""
import os
import warnings
import weakref

import torch

from . import _lib, ops

ENABLED = os.environ.get("A8_CUDA_GRAPHS", "1") != "0"
MAX_SHAPES = int(os.environ.get("A8_GRAPH_SHAPES", "2"))
CAPTURE_AFTER = int(os.environ.get("A8_GRAPH_AFTER", "2"))
WARMUP_ITERS = 2


def set_enabled(flag):
    """runtime switch (bench.py turns graphs off while it brackets single launches with CUDA events)"""
    global ENABLED
    ENABLED = bool(flag)


class GraphedSegment:
    def __init__(self, name):
        self.name = name
        self.entries = {}  # key -> (graphed callable, kernels per replay)
        self.seen = {}
        self.failed = set()
        self.live = {}  # key -> (weakref to the last replay's autograd node, [backward started?])
        self.warned = False

    def _key(self, inputs, params, extra):
        # the parameter part (address + requires_grad of ~200 tensors) is cached per params-tuple object and re-derived
        # when the probe (first / last address, count) changes: a module moved with .to() moves all of its storage
        probe = (id(params), len(params), params[0].data_ptr() if params else 0, params[-1].data_ptr() if params else 0,
                 params[-1].requires_grad if params else False)
        cached = self.__dict__.get("_pkey")
        if cached is None or cached[0] != probe:
            cached = self._pkey = (probe, tuple((p.data_ptr(), p.requires_grad) for p in params), params)
        return (tuple((tuple(t.shape), t.dtype, t.requires_grad) for t in inputs), cached[1], torch.is_grad_enabled(), extra,
                ops.grad_arena_serial())

    def run(self, fn, inputs, params, extra=(), clone_outputs=False):
        """fn(*inputs, *params) -> tensor or tuple of tensors, FUNCTIONAL in both (it must not reach parameters through
        module attributes).  `inputs` are the tensors whose VALUES change per call (fixed shapes); `params` the
        nn.Parameters the segment reads (their storage must not move).

        A replay returns ALIASES of the graph's static output buffers and keeps its activations in static buffers too:
        * clone_outputs=True (what the public module boundaries pass) hands the caller fresh tensors, so results kept
          across calls (`outs.append(encoder(x))`) are not overwritten by the next replay;
        * a second forward while the previous replay's autograd graph is still alive and has not started its backward
          would overwrite that graph's saved activations: such a call runs eagerly instead (one forward per backward is
          the replayed pattern; anything else is correct but slower)."""
        if not ENABLED or not inputs[0].is_cuda or torch.cuda.is_current_stream_capturing():
            return fn(*inputs, *params)
        if not torch.is_grad_enabled():
            # inference under no_grad is not the path these graphs exist for (make_graphed_callables captures a backward
            # for every parameter that requires grad, which a no_grad forward cannot provide): plain launches
            return fn(*inputs, *params)
        key = self._key(inputs, params, extra)
        ent = self.entries.get(key)
        if ent is None:
            n = self.seen.get(key, 0) + 1
            self.seen[key] = n
            if n < CAPTURE_AFTER or len(self.entries) >= MAX_SHAPES or key in self.failed:
                return fn(*inputs, *params)
            ent = self._capture(fn, inputs, params, key)
            if ent is None:
                return fn(*inputs, *params)
        graphed, n_kernels = ent
        prev = self.live.get(key)
        if prev is not None and not prev[1][0] and prev[0]() is not None:
            if not self.warned:
                self.warned = True
                warnings.warn(f"audio8_b200: segment '{self.name}' called again before the backward of its previous call: "
                              "running this call eagerly (a CUDA-graph replay would overwrite the saved activations)")
            return fn(*inputs, *params)
        _lib.load().a8_launch_count_add(n_kernels)
        out = graphed(*inputs, *params)
        outs = out if isinstance(out, (tuple, list)) else (out,)
        node = next((t.grad_fn for t in outs if isinstance(t, torch.Tensor) and t.grad_fn is not None), None)
        if node is not None:
            done = [False]

            def _started(_g, d=done):
                d[0] = True

            for t in outs:
                if isinstance(t, torch.Tensor) and t.requires_grad:
                    t.register_hook(_started)
            try:
                self.live[key] = (weakref.ref(node), done)
            except TypeError:  # a node type without weak references: fall back to the backward-started flag alone
                self.live[key] = ((lambda n=node: None), done)
        else:
            self.live.pop(key, None)
        if clone_outputs:
            out = tuple(t.clone() if isinstance(t, torch.Tensor) else t for t in outs) if isinstance(out, (tuple, list)) \
                else out.clone()
        return out

    def _capture(self, fn, inputs, params, key):
        lib = _lib.load()
        static_in = tuple(t.detach().clone().requires_grad_(t.requires_grad) for t in inputs)
        # The capture runs on ALIASES of the parameters (same storage, fresh autograd leaves).  The real parameters'
        # AccumulateGrad nodes belong to whatever stream first used them — normally the legacy default stream — and
        # the autograd engine would try to make that stream wait on the capturing one, which CUDA forbids.
        alias = tuple(p.detach().requires_grad_(p.requires_grad) for p in params)
        n0 = lib.a8_launch_count()
        # the warm-up iterations draw from the CUDA generator; put it back afterwards so that the step that captures sees
        # the same random stream (Gumbel noise, dropout seeds) as the eager step or the replay it stands in for
        rng = torch.cuda.get_rng_state(static_in[0].device)
        try:
            graphed = torch.cuda.make_graphed_callables(fn, static_in + alias, num_warmup_iters=WARMUP_ITERS,
                                                        allow_unused_input=True)
        except Exception as e:  # an un-capturable configuration runs eagerly; a broken capture must be loud once
            if os.environ.get("A8_GRAPH_STRICT"):
                raise
            self.failed.add(key)
            warnings.warn(f"audio8_b200: CUDA-graph capture of segment '{self.name}' failed, running eagerly: {e!r}")
            return None
        finally:
            torch.cuda.set_rng_state(rng, static_in[0].device)
        per_replay = (lib.a8_launch_count() - n0) // (WARMUP_ITERS + 1)
        ent = (graphed, int(per_replay))
        self.entries[key] = ent
        return ent
