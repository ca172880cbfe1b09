"""Drop-in for the hot-path classes of `audio8/wav2vec2.py` (reference: /root/reference/audio8/wav2vec2.py).

Same factories, constructor arguments, attribute names, return tuples and `state_dict` keys as the reference
(`create_model` :219, `create_acoustic_model` :262, `create_loss` :395, `ConvFeatureExtractionModel` :399,
`GumbelVectorQuantizer` :459, `AudioTransformerEncoder` :579, `Wav2Vec2Encoder` :649, `Wav2Vec2AcousticModel` :726,
`Wav2Vec2Model` :871, `Sampler` :955, `Wav2Vec2Loss` :371), so `pretrain.py` / `train.py` run unmodified when
`audio8.wav2vec2` resolves here.  The arithmetic runs in hand-written sm_100a kernels
(audio8_b200/csrc) through `functional.py`; host-side integer work (span masks, negative indices) repeats the
reference's numpy calls in the reference's order so that the draws are bit-identical.
"""
import contextlib
import math

import numpy as np
import torch
import torch.nn as nn

from . import functional as Fn
from .functional import BF16, F32
from .graphs import GraphedSegment

CONV_FEATURES = {
    16: [(512, 10, 5), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 2, 2), (512, 2, 2)],
    8: [(512, 10, 5), (512, 3, 2), (512, 3, 2), (512, 3, 2), (512, 2, 2), (512, 2, 2)],
}
START_TEMP = 2
END_TEMP = 0.5
TEMP_DECAY_FACTOR = 0.999995
XE_WGT = 0.1
DIVERSITY_WGT = 10


# --------------------------------------------------------------------------------------------------
# host-side integer work (numpy global RNG, same call order as the reference)
# --------------------------------------------------------------------------------------------------
def create_mask(shape, p_start=0.65, mask_length=10):
    """Span mask with every row subsampled to the batch-minimum count (reference wav2vec2.py:189-216).
    Draw order on numpy's global RNG: rand() once, choice() per row, choice() per over-long row."""
    bsz, T = shape
    mask = np.full((bsz, T), False)
    num_mask = int(p_start * T / float(mask_length) + np.random.rand())
    if num_mask == 0:
        return mask
    rows = []
    offs = np.arange(mask_length)[None, :]
    for i in range(bsz):
        span = mask_length
        if T - span <= num_mask:
            span = T - num_mask - 1
        starts = np.random.choice(T - span, num_mask, replace=False)
        idx = (starts[:, None] + offs).reshape(-1)
        mask[i, idx[idx < T]] = True
        rows.append(np.flatnonzero(mask[i]))  # == np.unique(idx[idx < T]) of the reference: sorted, no duplicates
    keep = min(len(r) for r in rows)
    for i, r in enumerate(rows):
        if len(r) > keep:
            mask[i] = False
            mask[i, np.random.choice(r, keep, replace=False)] = True
    return mask


class _PinnedRing:
    """Persistent pinned staging memory for the small per-step index uploads (mask rows, negative indices), used as a
    ring: a slice is reused only after the copy that last read it has completed (CUDA event per slice).  Neither a
    pageable source (the copy would make the host wait for all queued GPU work) nor per-call pinned allocations
    (cudaHostAlloc costs milliseconds) are acceptable inside the step."""

    def __init__(self, nbytes=8 << 20):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        self.np = self.buf.numpy()  # staging copies go through numpy: a torch CPU copy_ of this size wakes the OpenMP
        self.off = 0                # pool, whose workers then spin for their block time and preempt the enqueuing threads
        self.inflight = []  # (start, end, event)

    def upload(self, arr, device):
        """arr: contiguous numpy array -> device tensor of the same dtype/shape (asynchronous copy)"""
        n = arr.nbytes
        if n > self.buf.numel():
            self.__init__(2 * n)
        start = (self.off + 255) & ~255
        if start + n > self.buf.numel():
            start = 0
        end = start + n
        keep = []
        for (a, b, ev) in self.inflight:
            if a < end and start < b:
                ev.synchronize()
            elif not ev.query():
                keep.append((a, b, ev))
        self.inflight = keep
        np.copyto(self.np[start:end].view(arr.dtype).reshape(arr.shape), arr)
        stage = self.buf[start:end].view(_NP2TORCH[arr.dtype.type]).view(arr.shape)
        out = stage.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.inflight.append((start, end, ev))
        self.off = end
        return out


_NP2TORCH = {np.int32: torch.int32, np.int64: torch.int64, np.uint8: torch.uint8, np.bool_: torch.bool,
             np.float32: torch.float32}
_RING = {}


def _to_device(arr, device):
    """numpy -> device, asynchronously (through the pinned ring)"""
    arr = np.ascontiguousarray(arr)
    device = torch.device(device)
    if device.type != "cuda" or arr.size == 0 or arr.dtype.type not in _NP2TORCH:
        return torch.from_numpy(arr).to(device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ring = _RING.get(key)
    if ring is None:
        ring = _RING[key] = _PinnedRing()
    return ring.upload(arr, device)


def _mask_rows(mask_np, device):
    """flat row indices (b*T + t, row-major: the order boolean indexing produces) as int32 on the device"""
    idx = np.flatnonzero(mask_np.reshape(-1)).astype(np.int32)
    return _to_device(idx, device)


def conv_out_length(L, conv_features):
    """frames produced by the conv feature encoder for L samples (no padding: floor((L - k) / s) + 1 per layer)"""
    for (_, k, s) in conv_features:
        L = (L - k) // s + 1
    return L


class Sampler:
    """Negative sampling among the masked steps of the same utterance (reference wav2vec2.py:955-976)."""

    def __init__(self, n_negatives=100):
        self.n_negatives = n_negatives

    def indices(self, B, T):
        """[B, K*T] int64 numpy, already offset by b*T; identical draws to the reference's np.random.randint"""
        return self.indices32(B, T).astype(np.int64)

    def indices32(self, B, T):
        """same values as int32 (what the kernels read).  Only the randint call is the reference's; the never-the-
        positive shift and the per-utterance offset are done in place on the narrow copy (the int64 `where` of the
        reference costs as much host time as the draw itself)."""
        own_t = np.repeat(np.arange(T, dtype=np.int32), self.n_negatives)[None, :]
        idx = np.random.randint(0, T - 1, (B, self.n_negatives * T)).astype(np.int32)
        idx += idx >= own_t
        idx += (np.arange(B, dtype=np.int32) * T)[:, None]
        return idx

    def negatives(self, y):
        """reference-compatible API: returns (negs [K,B,T,C], neg_idxs [B,K*T]); the fused loss never calls this"""
        B, T, C = y.shape
        idx = torch.from_numpy(self.indices(B, T))
        negs = y.reshape(-1, C)[idx.view(-1).to(y.device)]
        return negs.view(B, T, self.n_negatives, C).permute(2, 0, 1, 3), idx


_DRAW_THREAD = __import__("os").environ.get("A8_HOST_DRAWS_THREAD", "1") != "0"
_ENCODER_LAST = __import__("os").environ.get("A8_ENCODER_LAST", "1") != "0"


def _rng_state_equal(a, b):
    return b is not None and a[0] == b[0] and a[2] == b[2] and a[3] == b[3] and a[4] == b[4] and np.array_equal(a[1], b[1])


class _HostDraws:
    """All numpy-RNG draws of one pre-training step, made on a helper thread in the reference's order — span mask
    (wav2vec2.py:189-216), one draw per transformer layer (eight_mile's LayerDrop test), negative indices
    (wav2vec2.py:967) — while the main thread keeps enqueuing GPU work.  numpy's legacy generator releases the GIL
    inside randint / permutation, so the ~3 ms of host time a step spends drawing no longer stalls the stream.
    The draws are the same calls on the same global generator in the same order: results stay bit-identical to the
    reference's as long as nothing else draws from np.random during the step (nothing in this package does)."""

    prefetch = False   # set by the trainer / bench: draw the NEXT step's numbers while the GPU runs the current step
    _pending = None    # the prefetched draws, if any

    @classmethod
    def get(cls, B, T, p_start, mask_length, n_layers, sampler):
        """the draws of the step that starts now.  With `prefetch` they were usually made during the previous step; they
        are used only if they are what the reference would draw now: same shapes AND numpy's global generator is in
        exactly the state the prefetch left it in (nobody reseeded it or drew from it since).  Otherwise the generator
        is put back where the prefetch found it (shape change) or left alone (somebody else touched it) and the
        numbers are drawn afresh — bit-identical to the reference either way."""
        pend, cls._pending = cls._pending, None
        if pend is not None:
            pend.ev_neg.wait()
            untouched = _rng_state_equal(np.random.get_state(), pend.state_after)
            if untouched and pend.error is None and pend.args[:5] == (B, T, p_start, mask_length, n_layers) and \
                    pend.args[5].n_negatives == sampler.n_negatives:
                return pend
            if untouched:
                np.random.set_state(pend.state_before)
        return cls(B, T, p_start, mask_length, n_layers, sampler)

    def consumed(self):
        """the step has taken its negatives: with `prefetch`, start drawing for the next step (same shapes assumed)"""
        if _HostDraws.prefetch and _DRAW_THREAD and _HostDraws._pending is None:
            _HostDraws._pending = _HostDraws(*self.args)

    def __init__(self, B, T, p_start, mask_length, n_layers, sampler=None):
        import threading
        self.args = (B, T, p_start, mask_length, n_layers, sampler)
        self.mask = self.layer_draws = self.neg = self.error = None
        self.state_before = np.random.get_state()
        self.state_after = None
        self.ev_mask, self.ev_neg = threading.Event(), threading.Event()
        if _DRAW_THREAD:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        else:  # A8_HOST_DRAWS_THREAD=0: same draws, made inline by the calling thread
            self._run()

    def _run(self):
        B, T, p_start, mask_length, n_layers, sampler = self.args
        try:
            self.mask = create_mask((B, T), p_start=p_start, mask_length=mask_length)
            self.layer_draws = [np.random.random() for _ in range(n_layers)]
        except BaseException as e:  # re-raised on the consumer's thread
            self.error = e
        self.ev_mask.set()
        try:
            if self.error is None and sampler is not None:
                self.neg = sampler.indices32(B, int(self.mask[0].sum()))
        except BaseException as e:
            self.error = e
        self.state_after = np.random.get_state()
        self.ev_neg.set()

    def time_mask(self):
        self.ev_mask.wait()
        if self.error is not None:
            raise self.error
        return self.mask, self.layer_draws

    def negatives(self):
        self.ev_neg.wait()
        if self.error is not None:
            raise self.error
        return self.neg

_DEVICE_DRAWS = [__import__("os").environ.get("A8_DEVICE_DRAWS", "0") == "1"]


def set_device_draws(flag):
    """Opt-in: draw the span mask and the negative indices of a pre-training step ON THE DEVICE (csrc/draws.cu, Philox)
    instead of with numpy's global generator on the host.  Same distributions as the reference, not the
    same numbers: the default (host) mode is the one that is bit-identical to the reference's draws.  In this mode the
    step uploads no index list, the draws live inside the CUDA-graph segments and change with torch's CUDA generator
    (`torch.manual_seed`).  Because the number of masked frames is then known only on the device, `forward` returns
    the latents PADDED: `y` is [B, R_max / B, C] (real rows first, flat over the batch, then zero rows) and
    `time_mask.a8_rows[-1]` holds the real count; `Wav2Vec2Loss` consumes exactly that."""
    _DEVICE_DRAWS[0] = bool(flag)


def set_prefetch_draws(flag):
    """Trainer-level switch: make the numpy draws of step i+1 (span mask, LayerDrop, negatives) on the helper thread while
    the GPU runs step i, instead of at the start of step i+1.  The numbers are the ones the reference would draw (same
    calls, same order, same global generator) provided nothing else draws from `np.random` between two steps; a reseed
    or a shape change between steps is detected and handled exactly (see _HostDraws.get)."""
    _HostDraws.prefetch = bool(flag)
    if not flag:
        _HostDraws._pending = None


# --------------------------------------------------------------------------------------------------
# small modules with reference-identical parameter names
# --------------------------------------------------------------------------------------------------
class _Linear(nn.Module):
    """nn.Linear parameters (weight [out,in], bias) evaluated by the tcgen05 GEMM"""

    def __init__(self, in_sz, out_sz):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_sz, in_sz))
        self.bias = nn.Parameter(torch.zeros(out_sz))
        nn.init.xavier_uniform_(self.weight)
        self._ops = {}  # bf16 operand copy of the weight, refreshed when the parameter changes (functional.operands)

    def refresh_operands(self):
        Fn.linear_operands(self._ops, self.weight, self.bias)

    def forward(self, x, out_f32=False):
        return Fn.linear(x, self.weight, self.bias, out_f32, self._ops)


class Dense(nn.Module):
    """eight_mile Dense: `.layer` is the Linear (keys `*.layer.weight`, reference wav2vec2.py:132-137)"""

    def __init__(self, insz, outsz):
        super().__init__()
        self.layer = _Linear(insz, outsz)
        self.output_dim = outsz

    def refresh_operands(self):
        self.layer.refresh_operands()

    def forward(self, x, out_f32=False):
        return self.layer(x, out_f32)


class _Conv1dParams(nn.Module):
    def __init__(self, cin, cout, k, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, k))
        nn.init.kaiming_uniform_(self.weight)
        if bias:
            self.bias = nn.Parameter(torch.zeros(cout))


class _Affine(nn.Module):
    def __init__(self, n):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(n))
        self.bias = nn.Parameter(torch.zeros(n))


class ConvFeatureExtractionModel(nn.Module):
    """Reference wav2vec2.py:399-456.  Keys: conv_layers.{i}.0.weight, conv_layers.0.2.{weight,bias} (GroupNorm)."""

    def __init__(self, conv_layers, dropout=0.0, conv_bias=False):
        super().__init__()
        if conv_bias or dropout != 0.0:
            raise NotImplementedError("the reference only ever uses conv_bias=False, dropout=0.0 (wav2vec2.py:403-404)")
        self.spec = [tuple(c) for c in conv_layers]
        self.conv_layers = nn.ModuleList()
        cin = 1
        for i, (dim, k, stride) in enumerate(self.spec):
            mods = [_Conv1dParams(cin, dim, k), nn.Identity()]
            if i == 0:
                mods.append(_Affine(dim))  # GroupNorm(dim, dim) affine parameters at index 2
            mods.append(nn.Identity())
            self.conv_layers.append(nn.ModuleList(mods))
            cin = dim
        self._ops = {}  # packed bf16 conv weights (forward and data-gradient layouts), refreshed when a weight changes

    def refresh_operands(self):
        Fn.conv_operands(self._ops, self.spec, [layer[0].weight for layer in self.conv_layers])

    def forward_channels_last(self, x):
        """[B,L] fp32 -> bf16 [B,T,C]"""
        gn = self.conv_layers[0][2]
        weights = [layer[0].weight for layer in self.conv_layers]
        self.refresh_operands()
        return Fn.ConvFeatureFn.apply(x, self.spec, self._ops, gn.weight, gn.bias, *weights)

    def forward(self, x):
        """reference layout: [B,C,T] fp32"""
        return self.forward_channels_last(x).float().transpose(1, 2)


class GumbelVectorQuantizer(nn.Module):
    """Reference wav2vec2.py:459-576 (keys: vars [1,G*V,vd], weight_proj.{weight,bias})."""

    def __init__(self, dim, num_vars, min_temperature, max_temperature, temperature_decay, num_groups, vq_dim):
        super().__init__()
        self.num_groups = num_groups
        self.input_dim = dim
        self.num_vars = num_vars
        assert vq_dim % num_groups == 0, f"dim {vq_dim} must be divisible by groups {num_groups} for concatenation"
        var_dim = vq_dim // num_groups
        self.vars = nn.Parameter(torch.FloatTensor(1, num_groups * num_vars, var_dim))
        nn.init.uniform_(self.vars)
        self.weight_proj = _Linear(self.input_dim, num_groups * num_vars)
        nn.init.normal_(self.weight_proj.weight, mean=0, std=1)
        nn.init.zeros_(self.weight_proj.bias)
        self.max_temperature = max_temperature
        self.min_temperature = min_temperature
        self.temperature_decay = temperature_decay
        self.curr_temperature = self.max_temperature
        self.codebook_indices = None
        self.noise_override = None  # parity tests: Gumbel noise [B*T*G, V] drawn like F.gumbel_softmax does
        self.last_indices = None
        self.keep_logits = False  # parity tests: keep a copy of the fp32 logits [B*T, G*V] in `last_logits`
        self.last_logits = None
        self._ops = {}  # fp32 / bf16x3-split / bf16 copies of weight_proj and the codebook (functional.operands)

    def refresh_operands(self):
        Fn.quantizer_operands(self._ops, self.weight_proj.weight, self.vars)

    def set_num_updates(self, num_updates):
        self.curr_temperature = max(self.max_temperature * self.temperature_decay ** num_updates, self.min_temperature)

    def forward(self, x, n_valid=None, _params=None):
        """x [B, Tm, C].  n_valid (int32 device scalar, optional): only the first n_valid of the B*Tm rows exist (the
        rest pads a worst-case-length row list).  _params: (weight, bias, vars) aliases when run inside a CUDA-graph
        capture (graphs.py hands the segment functional copies of the parameters)."""
        B, T, _ = x.shape
        w, b, v = _params if _params is not None else (self.weight_proj.weight, self.weight_proj.bias, self.vars)
        noise = None
        if self.training:
            noise = self.noise_override
            if noise is None:  # what F.gumbel_softmax draws (wav2vec2.py:557)
                n = B * T * self.num_groups
                noise = Fn.ops.bucketed_empty((n, self.num_vars), F32, x.device).exponential_().log_().neg_()
            else:
                noise = noise.to(device=x.device, dtype=F32).contiguous()
                n = B * T * self.num_groups
                if noise.shape[0] < n:  # parity tests supply noise for the valid rows only: pad rows are never read
                    noise = torch.cat([noise, noise.new_zeros(n - noise.shape[0], noise.shape[1])])
        Fn.QuantizerFn.keep_logits = self.keep_logits
        q, ppl, kidx = Fn.QuantizerFn.apply(x, w, b, v, self.num_groups, self.curr_temperature, noise, n_valid, self._ops)
        self.last_indices = kidx
        if self.keep_logits:
            self.last_logits = Fn.QuantizerFn.last_logits.detach().clone()
            Fn.QuantizerFn.keep_logits, Fn.QuantizerFn.last_logits = False, None
        return q, ppl


class _PosConv(nn.Module):
    """Conv1DSame(...).conv = Sequential(pad, conv, act) with weight_norm(dim=2) on conv[1]
    (keys pos_conv.conv.1.{bias,weight_g,weight_v}, reference wav2vec2.py:140-142, 600-609)"""

    def __init__(self, d_model, k, groups, std):
        super().__init__()
        conv = nn.Module()
        conv.bias = nn.Parameter(torch.zeros(d_model))
        v = torch.empty(d_model, d_model // groups, k).normal_(0, std)
        conv.weight_g = nn.Parameter(v.norm(2, dim=(0, 1), keepdim=True))
        conv.weight_v = nn.Parameter(v)
        self.conv = nn.ModuleList([nn.Identity(), conv, nn.Identity()])

    def weight(self):
        """the effective (weight-normed) kernel, for inspection; the training path fuses this into a pack kernel"""
        c = self.conv[1]
        return c.weight_g * c.weight_v / c.weight_v.norm(2, dim=(0, 1), keepdim=True)


class _MHAParams(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.w_Q, self.w_K, self.w_V, self.w_O = Dense(d, d), Dense(d, d), Dense(d, d), Dense(d, d)


class _LayerParams(nn.Module):
    """parameter container with eight_mile TransformerEncoder's names: self_attn.w_{Q,K,V,O}.layer, ffn.{0,3}.layer,
    ln1, ln2 (reference wav2vec2.py:110-126)"""

    def __init__(self, d, d_ff):
        super().__init__()
        self.self_attn = _MHAParams(d)
        self.ffn = nn.ModuleList([Dense(d, d_ff), nn.Identity(), nn.Identity(), Dense(d_ff, d)])
        self.ln1 = _Affine(d)
        self.ln2 = _Affine(d)

    def flat(self):
        a = self.self_attn
        return [a.w_Q.layer.weight, a.w_Q.layer.bias, a.w_K.layer.weight, a.w_K.layer.bias, a.w_V.layer.weight,
                a.w_V.layer.bias, a.w_O.layer.weight, a.w_O.layer.bias, self.ln2.weight, self.ln2.bias,
                self.ffn[0].layer.weight, self.ffn[0].layer.bias, self.ffn[3].layer.weight, self.ffn[3].layer.bias,
                self.ln1.weight, self.ln1.bias]


class _Stack(nn.Module):
    def __init__(self, d, d_ff, layers):
        super().__init__()
        self.encoders = nn.ModuleList([_LayerParams(d, d_ff) for _ in range(layers)])


def _encoder_backward_done(grad):
    Fn.ops.encoder_backward_done()
    return grad


class AudioTransformerEncoder(nn.Module):
    """Reference wav2vec2.py:579-646: positional conv (k=128, groups=16, weight-normed) + GELU, residual, LayerNorm,
    dropout, then a post-LN transformer stack (eight_mile TransformerEncoderStack, layer_norms_after=True)."""

    def __init__(self, num_heads, d_model, pdrop, layers=1, activation="gelu", d_ff=None, conv_pos_kernel=128,
                 conv_groups=16, layer_drop=0.0, **kwargs):
        super().__init__()
        if activation != "gelu":
            raise NotImplementedError("the reference always uses gelu here (wav2vec2.py:618)")
        self.d_model = d_model
        self.num_heads = num_heads
        self.pdrop = pdrop
        self.conv_pos_kernel = conv_pos_kernel
        self.conv_groups = conv_groups
        self.layer_drop = layer_drop
        std = math.sqrt((4 * (1.0 - pdrop)) / (conv_pos_kernel * d_model))
        self.pos_conv = _PosConv(d_model, conv_pos_kernel, conv_groups, std)
        self.transformer = _Stack(d_model, d_ff if d_ff else 4 * d_model, layers)
        self.ln = _Affine(d_model)
        self._arena = {}  # per stack slice: persistent bf16 operand buffers + device pointer table (values rewritten every call)
        self._graph = GraphedSegment("transformer encoder (front + lower layers)")
        self._graph2 = GraphedSegment("transformer encoder (upper layers)")
        # measured at N=2 (round 1): the split neither helps nor hurts the step time, so it stays off by default
        self.split_min_layers = int(__import__("os").environ.get("A8_ENCODER_SPLIT_MIN", "1000"))

    def forward(self, x, pad_mask=None):
        return self.extract_features(x, pad_mask)

    def extract_features(self, x, pad_mask=None, _layer_draws=None, _internal=False):
        """x [B,T,D] (bf16 or fp32), pad_mask bool [B,T] (True = valid) or None -> bf16 [B,T,D].
        The result is a fresh tensor (a CUDA-graph replay's static output buffer is cloned at this public boundary);
        `_internal` callers that consume the result within the same step skip the copy."""
        if x.requires_grad and Fn.ops.grad_arena_active():
            # data-parallel wrapper (parallel.py): when the gradient w.r.t. the encoder's input exists, every layer has
            # written its gradients into the arena and their all-reduce can start under the rest of backward
            x.register_hook(_encoder_backward_done)
        n = len(self.transformer.encoders)
        active = []
        for i in range(n):  # eight_mile draws one numpy random per layer, even with layer_drop == 0
            pdrop = np.random.random() if _layer_draws is None else _layer_draws[i]
            active.append((not self.training) or pdrop >= self.layer_drop)
        row_keep = None
        if pad_mask is not None:
            row_keep = pad_mask.to(device=x.device, dtype=torch.uint8).contiguous()
        pc = self.pos_conv.conv[1]
        front_params = (pc.weight_g, pc.weight_v, pc.bias, self.ln.weight, self.ln.bias)
        # The stack runs as one or two autograd nodes / CUDA-graph segments.  Two (split in the middle) when it is deep
        # enough: under DistributedDataParallel the gradients of the upper half then leave their node — and their
        # buckets start to all-reduce — while the lower half's backward is still running.
        cuts = [0, n // 2, n] if (n >= self.split_min_layers and all(active)) else [0, n]
        h = x
        for part in range(len(cuts) - 1):
            lo, hi = cuts[part], cuts[part + 1]
            cfg = dict(num_heads=self.num_heads, groups=self.conv_groups, pdrop=self.pdrop, training=self.training,
                       active=active[lo:hi], arena=self._arena.setdefault(part, {}), front=(part == 0))
            params = self._part_params(part, lo, hi, front_params)
            # bf16 / packed operand copies: refreshed here, eagerly, only if a parameter changed since the last call
            Fn.encoder_operands(cfg["arena"], pc.weight_g, pc.weight_v, params[5:] if part == 0 else params,
                                Fn.EncoderFn.PER_LAYER, front=(part == 0))
            h = self._run_part(part, h, cfg, row_keep, params, clone=(not _internal) and part == len(cuts) - 2)
        return h

    def _part_params(self, part, lo, hi, front_params):
        """the flat parameter tuple of a stack slice, built once (192 attribute look-ups through nn.Module per call
        otherwise); the same tuple OBJECT is handed to the graph segment so that its key cache hits"""
        cache = self.__dict__.setdefault("_pp_cache", {})
        probe = (lo, hi, id(self.transformer.encoders[lo].self_attn.w_Q.layer.weight), id(front_params[0]))
        ent = cache.get(part)
        if ent is None or ent[0] != probe:
            flat = []
            for layer in self.transformer.encoders[lo:hi]:
                flat += layer.flat()
            ent = cache[part] = (probe, (*front_params, *flat) if part == 0 else tuple(flat))
        return ent[1]

    def _run_part(self, part, x, cfg, row_keep, params, clone=False):
        nf = 5 if cfg["front"] else 0

        def call(x_, rk, ps):
            fr = ps[:nf] if nf else (None,) * 5
            return Fn.EncoderFn.apply(x_, cfg, rk, *fr, *ps[nf:])

        if not all(cfg["active"]):  # LayerDrop changes the launch sequence per step: eager
            return call(x, row_keep, params)
        # static shapes: replay a captured CUDA graph once this (shape, mode) has been seen before (graphs.py)
        seg = self._graph if part == 0 else self._graph2
        if row_keep is None:
            return seg.run(lambda x_, *ps: call(x_, None, ps), (x,), params,
                           extra=(self.training, self.pdrop, part, Fn.ops.grad_arena_active()), clone_outputs=clone)
        return seg.run(lambda x_, rk, *ps: call(x_, rk, ps), (x, row_keep), params,
                       extra=(self.training, self.pdrop, part, Fn.ops.grad_arena_active()), clone_outputs=clone)


class Wav2Vec2Encoder(nn.Module):
    """Reference wav2vec2.py:649-723 (fine-tuning encoder: pad handling, time + channel masking when training)."""

    def __init__(self, conv_features=CONV_FEATURES[16], d_model=768, num_heads=12, num_layers=12, dropout=0.1, d_ff=None,
                 dropout_input=0.1, dropout_features=0.0, timestep_masking=0.5, channel_masking=0.1,
                 timestep_mask_len=10, channel_mask_len=64, layer_drop=0.0, freeze_fx=True):
        super().__init__()
        fx_dsz = conv_features[-1][0]
        self.layer_norm = _Affine(fx_dsz)
        self.dropout_input_p = dropout_input
        self.dropout_features_p = dropout_features
        self.feature_extractor = ConvFeatureExtractionModel(conv_features)
        self.proj_to_input = Dense(fx_dsz, d_model)
        self.encoder = AudioTransformerEncoder(num_heads, d_model, dropout, num_layers, d_ff=d_ff, layer_drop=layer_drop)
        self.mask_emb = nn.Parameter(torch.FloatTensor(d_model).uniform_())
        self.timestep_masking = timestep_masking
        self.channel_masking = channel_masking
        self.timestep_mask_len = timestep_mask_len
        self.channel_mask_len = channel_mask_len
        self.output_dim = d_model
        self.freeze_fx = freeze_fx

    def forward(self, x, pad_mask=None):
        with torch.no_grad() if self.freeze_fx else contextlib.ExitStack():
            fx = self.feature_extractor.forward_channels_last(x)
        features = Fn.layer_norm(fx, self.layer_norm.weight, self.layer_norm.bias, 1e-5)
        B, T, _ = features.shape
        if pad_mask is not None:  # reference :703-708
            extra = pad_mask.size(1) % T
            if extra > 0:
                pad_mask = pad_mask[:, :-extra]
            pad_mask = pad_mask.view(pad_mask.size(0), T, -1).all(-1)
        features = self.proj_to_input(features)
        C = features.shape[-1]
        features = Fn.dropout(features, self.dropout_input_p, self.training)
        if self.training and self.timestep_masking > 0.0:
            time_mask = create_mask((B, T), p_start=self.timestep_masking, mask_length=self.timestep_mask_len)
            if time_mask.any():  # num_mask == 0 on very short utterances: `features[time_mask] = ...` is a no-op (:717)
                features = Fn.RowsSetFn.apply(features, _mask_rows(time_mask, x.device), self.mask_emb)
        if self.training and self.channel_masking > 0.0:
            channel_mask = create_mask((B, C), p_start=self.channel_masking, mask_length=self.channel_mask_len)
            if channel_mask.any():
                cz = _to_device(channel_mask.astype(np.uint8), x.device)
                features = Fn.MaskApplyFn.apply(features, None, cz)
        out = self.encoder(features, pad_mask)
        return out, pad_mask


class Wav2Vec2AcousticModel(nn.Module):
    """Reference wav2vec2.py:726-770: encoder + linear head + log_softmax; `freeze` gates the encoder's gradients."""

    def __init__(self, num_labels, conv_features=CONV_FEATURES[16], d_model=768, num_heads=12, num_layers=12,
                 dropout=0.1, d_ff=None, dropout_input=0.0, dropout_features=0.0, timestep_masking=0.5,
                 channel_masking=0.1, timestep_mask_len=10, channel_mask_len=64, layer_drop=0.0, freeze_fx=True):
        super().__init__()
        self.encoder = Wav2Vec2Encoder(conv_features, d_model, num_heads, num_layers, dropout, d_ff, dropout_input,
                                       dropout_features, timestep_masking, channel_masking, timestep_mask_len,
                                       channel_mask_len, layer_drop, freeze_fx=freeze_fx)
        self.proj = _Linear(d_model, num_labels)
        self.freeze = True

    def forward(self, x, pad_mask=None):
        with torch.no_grad() if self.freeze else contextlib.ExitStack():
            encoded, pad_mask = self.encoder(x, pad_mask)
        logits = self.proj(encoded, out_f32=True)
        lp = Fn.LogSoftmaxFn.apply(logits)
        # the CTC loss (audio8_b200.ctc) recognises this tensor and works from the logits: fused log_softmax + CTC
        lp.a8_logits = logits if logits.is_contiguous() else None
        return lp, pad_mask


class Wav2Vec2Model(nn.Module):
    """Reference wav2vec2.py:871-952: contrastive pre-training model.  Returns (x [B,T,final], y [B,Tm,final],
    vq perplexity (scalar), time_mask bool [B,T]) like the reference; the mask's row indices ride along on the
    returned tensor (`time_mask.a8_rows`) so the loss needs no nonzero / host sync."""

    def __init__(self, conv_features=CONV_FEATURES[16], num_vq_vars=320, start_temp=START_TEMP, end_temp=END_TEMP,
                 temp_decay_factor=TEMP_DECAY_FACTOR, num_vq_groups=2, d_model=768, num_heads=12, num_layers=12,
                 dropout=0.1, d_ff=None, final_dim=256, dropout_input=0.1, dropout_features=0.1, timestep_masking=0.65,
                 channel_masking=0.0, timestep_mask_len=10, channel_mask_len=64, layer_drop=0.0):
        super().__init__()
        fx_dsz = conv_features[-1][0]
        self.layer_norm = _Affine(fx_dsz)
        self.dropout_input_p = dropout_input
        self.dropout_features_p = dropout_features
        self.feature_extractor = ConvFeatureExtractionModel(conv_features)
        self.proj_to_input = Dense(fx_dsz, d_model)
        # the reference passes (start, end, decay) into (min, max, decay) slots: tau == end_temp
        self.quantizer = GumbelVectorQuantizer(fx_dsz, num_vq_vars, start_temp, end_temp, temp_decay_factor,
                                               num_vq_groups, final_dim)
        self.encoder = AudioTransformerEncoder(num_heads, d_model, dropout, num_layers, d_ff=d_ff, layer_drop=layer_drop)
        self.project_q = Dense(final_dim, final_dim)
        self.final_proj = Dense(d_model, final_dim)
        self.timestep_masking = timestep_masking
        self.channel_masking = channel_masking
        self.timestep_mask_len = timestep_mask_len
        self.channel_mask_len = channel_mask_len
        self.mask_emb = nn.Parameter(torch.FloatTensor(d_model).uniform_())
        self._front_graph = GraphedSegment("conv feature encoder + LayerNorm + input projection + time mask")
        self._branch_graph = GraphedSegment("quantizer branch (gather, Gumbel VQ, project_q)")

    def _front_params(self):
        fe = self.feature_extractor
        gn = fe.conv_layers[0][2]
        return (gn.weight, gn.bias, *[layer[0].weight for layer in fe.conv_layers], self.layer_norm.weight,
                self.layer_norm.bias, self.proj_to_input.layer.weight, self.proj_to_input.layer.bias, self.mask_emb)

    def _branch_params(self):
        q = self.quantizer
        return (q.weight_proj.weight, q.weight_proj.bias, q.vars, self.project_q.layer.weight, self.project_q.layer.bias)

    def _front(self, x, rows, gn_w, gn_b, *rest):
        """audio -> (projected, dropped-out features with the mask embedding at the masked frames, bf16 [B,T,D];
        un-projected LayerNorm output fp32 [B,T,512]); reference :929-939.  Functional in the parameters (order of
        `_front_params`) so that a CUDA-graph capture can run it on aliases of them (graphs.py)."""
        n = len(self.feature_extractor.spec)
        conv_w, (ln_w, ln_b, pw, pb, mask_emb) = rest[:n], rest[n:]
        fx = Fn.ConvFeatureFn.apply(x, self.feature_extractor.spec, self.feature_extractor._ops, gn_w, gn_b, *conv_w)
        features, unmasked = Fn.layer_norm(fx, ln_w, ln_b, 1e-5, want_f32=True)
        features = Fn.linear(features, pw, pb, cache=self.proj_to_input.layer._ops)
        features = Fn.dropout(features, self.dropout_input_p, self.training)
        features = Fn.RowsSetFn.apply(features, rows[:-1], mask_emb)
        return features, unmasked

    def _front_dev(self, x, gn_w, gn_b, *rest):
        """`_front` with the span mask drawn on the device inside the same segment (set_device_draws)"""
        B, T = x.shape[0], conv_out_length(x.shape[1], self.feature_extractor.spec)
        rows, mask = Fn.span_mask_draw(x, B, T, self.timestep_masking, self.timestep_mask_len, self.max_masked_rows(B, T))
        features, unmasked = self._front(x, rows, gn_w, gn_b, *rest)
        return features, unmasked, rows, mask

    def _branch(self, unmasked, rows, wq, bq, vars_, pw, pb):
        """quantizer branch (reference :946-950): masked frames of the un-projected features -> dropout -> Gumbel VQ ->
        project_q.  Works on the padded row list (rows[:-1], -1 = padding; rows[-1] = number of real rows)."""
        B, _, C = unmasked.shape
        y = Fn.RowsGatherFn.apply(unmasked, rows[:-1]).view(B, -1, C)
        y = Fn.dropout(y, self.dropout_features_p, self.training)
        q, vq_probs = self.quantizer(y, n_valid=rows[-1:], _params=(wq, bq, vars_))
        return Fn.linear(q, pw, pb, out_f32=True, cache=self.project_q.layer._ops), vq_probs

    def set_num_updates(self, s):
        self.quantizer.set_num_updates(s)

    def request_host_draws(self, sampler):
        """the loss announces that the next forward() belongs to a training step that will also need negative indices:
        forward() then makes all of the step's numpy draws on a helper thread (see _HostDraws), started right after
        the first GPU work of the step has been enqueued, and leaves the handle in `self._host_draws`"""
        self._host_draw_request = sampler

    def max_masked_rows(self, B, T):
        """worst-case number of masked frames of a batch: every span of every row lands without overlap"""
        n_spans = int(self.timestep_masking * T / float(self.timestep_mask_len)) + 1
        return B * min(T, n_spans * self.timestep_mask_len)

    def forward(self, x):
        sampler = self.__dict__.pop("_host_draw_request", None)
        self._host_draws = draws = None
        B = x.shape[0]
        T = conv_out_length(x.shape[1], self.feature_extractor.spec)
        if self.channel_masking > 0.0:
            raise NotImplementedError("channel masking in pre-training is broken in the reference (wav2vec2.py:943)")
        # ---- the step's numpy draws (mask, LayerDrop, negatives) in the reference's order; they depend on shapes only, so
        # they are made BEFORE any GPU work (prefetched during the previous step when the caller allows it: _HostDraws)
        layer_draws = None
        if _DEVICE_DRAWS[0]:
            return self._forward_device_draws(x, B, T)
        if sampler is not None:
            self._host_draws = draws = _HostDraws.get(B, T, self.timestep_masking, self.timestep_mask_len,
                                                      len(self.encoder.transformer.encoders), sampler)
            time_mask, layer_draws = draws.time_mask()
        else:
            time_mask = create_mask((B, T), p_start=self.timestep_masking, mask_length=self.timestep_mask_len)
        # masked rows as a worst-case-length list (static shapes: every segment below replays as a CUDA graph):
        # [flat row indices, -1 padding ..., number of real rows]
        idx = np.flatnonzero(time_mask.reshape(-1)).astype(np.int32)
        R_max = max(self.max_masked_rows(B, T), idx.size)
        buf = np.full(R_max + 1, -1, dtype=np.int32)
        buf[:idx.size] = idx
        buf[R_max] = idx.size
        rows = _to_device(buf, x.device)
        Fn.ops.set_dynamic_rows(R_max, R_max)
        # operand copies of the parameters (packed conv weights, bf16 projections, quantizer splits): refreshed eagerly and
        # only when a parameter changed since the last call, so the graph segments below contain no re-layout kernels
        for m in (self.feature_extractor, self.proj_to_input, self.quantizer, self.project_q, self.final_proj):
            m.refresh_operands()
        eager = self.quantizer.noise_override is not None or self.quantizer.keep_logits  # parity-test hooks: no replay
        arena = Fn.ops.grad_arena_active()
        features, unmasked = self._front_graph.run(self._front, (x, rows), self._front_params(),
                                                   extra=(self.training, self.dropout_input_p, arena))
        branch = lambda: (self._branch(unmasked, rows, *self._branch_params()) if eager else
                          self._branch_graph.run(self._branch, (unmasked, rows), self._branch_params(),
                                                 extra=(self.training, self.dropout_features_p, arena,
                                                        self.quantizer.curr_temperature)))
        if _ENCODER_LAST and not arena:
            # The quantizer branch is built BEFORE the encoder: autograd runs ready nodes in reverse creation order, so in
            # backward the encoder's (long, graph-replayed) backward is enqueued right after final_proj's.
            y_pad, vq_probs = branch()
            enc = self.encoder.extract_features(features, None, layer_draws, _internal=True)
        else:  # the encoder first.  Under the data-parallel wrapper this is the order that lets autograd finish the
            # quantizer branch's backward BEFORE the encoder's, so its gradients join the early all-reduce (parallel.py)
            enc = self.encoder.extract_features(features, None, layer_draws, _internal=True)
            y_pad, vq_probs = branch()
        xo = self.final_proj(enc, out_f32=True)
        qz = self.quantizer  # inspection hooks: the real rows only
        if qz.last_indices is not None and qz.last_indices.numel() == R_max * qz.num_groups:
            qz.last_indices = qz.last_indices[:idx.size * qz.num_groups]
        if qz.last_logits is not None and qz.last_logits.shape[0] == R_max:
            qz.last_logits = qz.last_logits[:idx.size]
        mask_t = _to_device(time_mask, x.device)
        mask_t.a8_rows = rows       # padded row list + count: the loss needs no nonzero / host sync
        mask_t.a8_ypad = y_pad      # the padded latents the loss kernels work on (gradients flow into this tensor)
        Tm = idx.size // B
        y = y_pad.view(-1, y_pad.shape[-1])[:idx.size].view(B, Tm, -1)
        return xo, y, vq_probs, mask_t


def _forward_device_draws(self, x, B, T):
    """Wav2Vec2Model.forward with device-side draws: no numpy draw, no index upload, nothing data dependent on the host"""
    R_max = self.max_masked_rows(B, T)
    Fn.ops.set_dynamic_rows(R_max, R_max)
    for m in (self.feature_extractor, self.proj_to_input, self.quantizer, self.project_q, self.final_proj):
        m.refresh_operands()
    eager = self.quantizer.noise_override is not None or self.quantizer.keep_logits
    arena = Fn.ops.grad_arena_active()
    features, unmasked, rows, mask = self._front_graph.run(
        self._front_dev, (x,), self._front_params(),
        extra=(self.training, self.dropout_input_p, arena, "device draws", self.timestep_masking, self.timestep_mask_len))
    branch = lambda: (self._branch(unmasked, rows, *self._branch_params()) if eager else
                      self._branch_graph.run(self._branch, (unmasked, rows), self._branch_params(),
                                             extra=(self.training, self.dropout_features_p, arena,
                                                    self.quantizer.curr_temperature)))
    if _ENCODER_LAST and not arena:
        y_pad, vq_probs = branch()
        enc = self.encoder.extract_features(features, None, None, _internal=True)
    else:
        enc = self.encoder.extract_features(features, None, None, _internal=True)
        y_pad, vq_probs = branch()
    xo = self.final_proj(enc, out_f32=True)
    mask_t = mask.view(torch.bool)
    mask_t.a8_rows = rows
    mask_t.a8_ypad = y_pad
    mask_t.a8_device_draws = True
    return xo, y_pad.view(B, R_max // B, -1), vq_probs, mask_t


Wav2Vec2Model._forward_device_draws = _forward_device_draws


def _flush_feed():
    """the forward of the step is enqueued: let the input feed post its next H2D copy now (feed.flush_deferred)"""
    from . import feed
    if feed._DEFERRED:
        feed.flush_deferred()


class Wav2Vec2Loss(nn.Module):
    """Reference wav2vec2.py:371-392: 0.1 * CE(cos-sim logits over [positive | K negatives]) + 10 * (n_vars - ppl) / n_vars."""

    def __init__(self, n_vars, n_negatives=100):
        super().__init__()
        self.n_vars = n_vars
        self.sample = Sampler(n_negatives)
        self.last_neg_idx = None

    def __call__(self, model, features):
        inner = getattr(model, "module", model)  # DistributedDataParallel wraps the model (pretrain.py:158)
        draws = None
        if isinstance(inner, Wav2Vec2Model):
            inner.request_host_draws(self.sample)
        try:
            outputs, latents, gs_probs, time_mask = model(features)
        finally:
            if isinstance(inner, Wav2Vec2Model):
                inner.__dict__.pop("_host_draw_request", None)  # never leave a stale request behind a failed forward
                draws = inner.__dict__.pop("_host_draws", None)
        B, Tm, C = latents.shape
        rows = getattr(time_mask, "a8_rows", None)
        y_pad = getattr(time_mask, "a8_ypad", None)
        if rows is None or y_pad is None:  # a model that is not ours: plain (dynamic-shape) path
            rows = torch.cat([torch.nonzero(time_mask.reshape(-1)).reshape(-1).int(),
                              torch.tensor([B * Tm], dtype=torch.int32, device=outputs.device)])
            y_pad = latents.reshape(B * Tm, C)
        R_max = rows.numel() - 1
        if getattr(time_mask, "a8_device_draws", False):  # negatives drawn on the device inside the loss segment
            K = self.sample.n_negatives
            y2 = y_pad.reshape(R_max, C)
            self.last_rows = rows

            def fn(o, y_, g, r):
                # (device tensors; under a CUDA-graph replay they are the graph's static buffers, refreshed by the replay)
                self.last_neg_idx = idx = Fn.negatives_draw(r, B, K)
                return self._loss(o, y_, g, r, idx)

            if gs_probs.requires_grad and outputs.requires_grad:
                graph = self.__dict__.setdefault("_graph", GraphedSegment("contrastive loss (gather, cosine logits, CE)"))
                loss = graph.run(fn, (outputs, y2, gs_probs, rows), (), extra=(self.n_vars, K, B, "device draws"))
            else:
                loss = fn(outputs, y2, gs_probs, rows)
            _flush_feed()
            return loss
        # numpy draws, bit-exact with the reference's Sampler (already made on the helper thread when `draws`)
        neg = draws.negatives() if draws is not None else self.sample.indices32(B, Tm)
        assert neg.shape == (B, self.sample.n_negatives * Tm)
        self.last_neg_idx = neg
        K = self.sample.n_negatives
        if R_max > B * Tm:  # pad to the static length (the kernels never read the padding rows' candidates)
            pad = np.zeros(R_max * K, dtype=np.int32)
            pad[:neg.size] = neg.reshape(-1)
            neg = pad
        idx = _to_device(neg.reshape(-1), outputs.device)
        if draws is not None:
            draws.consumed()
        y2 = y_pad.reshape(R_max, C)
        graph = self.__dict__.setdefault("_graph", GraphedSegment("contrastive loss (gather, cosine logits, CE)"))
        if gs_probs.requires_grad and outputs.requires_grad:
            loss = graph.run(self._loss, (outputs, y2, gs_probs, rows, idx), (), extra=(self.n_vars, K))
        else:
            loss = self._loss(outputs, y2, gs_probs, rows, idx)
        _flush_feed()
        return loss

    def _loss(self, outputs, y2, gs_probs, rows, idx):
        xm = Fn.RowsGatherFn.apply(outputs, rows[:-1])  # outputs[time_mask] -> [R_max, C], zero rows for the padding
        loss, _ = Fn.ContrastiveFn.apply(xm, y2, idx, gs_probs, self.n_vars, XE_WGT, DIVERSITY_WGT, rows[-1:])
        return loss


def create_loss(n_vars, n_negatives):
    return Wav2Vec2Loss(n_vars, n_negatives)


def create_model(sample_rate=16, num_vq_vars=320, num_vq_groups=2, d_model=768, num_heads=12, num_layers=12,
                 dropout=0.1, d_ff=None, final_dim=256, dropout_input=0.1, dropout_features=0.1, timestep_masking=0.65,
                 channel_masking=0.0, timestep_mask_len=10, channel_mask_len=64, layer_drop=0.0, **kwargs):
    """Same signature as the reference factory (wav2vec2.py:219-259); extra kwargs are swallowed like there."""
    return Wav2Vec2Model(CONV_FEATURES[sample_rate], num_vq_vars, START_TEMP, END_TEMP, TEMP_DECAY_FACTOR,
                         num_vq_groups, d_model, num_heads, num_layers, dropout, d_ff, final_dim, dropout_input,
                         dropout_features, timestep_masking, channel_masking, timestep_mask_len, channel_mask_len,
                         layer_drop)


def create_acoustic_model(num_labels, sample_rate=16, d_model=768, num_heads=12, num_layers=12, dropout=0.1, d_ff=None,
                          dropout_input=0.0, timestep_masking=0.5, channel_masking=0.1, timestep_mask_len=10,
                          channel_mask_len=64, layer_drop=0.0, freeze_fx=True, **kwargs):
    """Same signature as the reference factory (wav2vec2.py:262-296)."""
    return Wav2Vec2AcousticModel(num_labels, CONV_FEATURES[sample_rate], d_model, num_heads, num_layers, dropout, d_ff,
                                 dropout_input, 0.0, timestep_masking, channel_masking, timestep_mask_len,
                                 channel_mask_len, layer_drop, freeze_fx)


# --------------------------------------------------------------------------------------------------
# fairseq checkpoint import (reference wav2vec2.py:36-186): same function names, same key mapping
# --------------------------------------------------------------------------------------------------
_FAIRSEQ_LAYER = (("self_attn.k_proj", "self_attn.w_K.layer"), ("self_attn.v_proj", "self_attn.w_V.layer"),
                  ("self_attn.q_proj", "self_attn.w_Q.layer"), ("self_attn.out_proj", "self_attn.w_O.layer"),
                  ("self_attn_layer_norm", "ln2"), ("fc1", "ffn.0.layer"), ("fc2", "ffn.3.layer"), ("final_layer_norm", "ln1"))


def fairseq_key_map(num_layers, ctc=False, sr=16):
    """{fairseq key: audio8 key} for a checkpoint with `num_layers` transformer layers: what the reference's
    `convert_keys` renames (W2V_MAP for a pre-training checkpoint, W2V_CTC_MAP[sr] for a fine-tuned one,
    wav2vec2.py:36-151); every other key keeps its name."""
    src, dst = ("w2v_encoder.w2v_model.", "encoder.") if ctc else ("", "")
    m = {}
    for i in range(num_layers):
        for fs, a8 in _FAIRSEQ_LAYER:
            for leaf in ("weight", "bias"):
                m[f"{src}encoder.layers.{i}.{fs}.{leaf}"] = f"{dst}encoder.transformer.encoders.{i}.{a8}.{leaf}"
    for leaf in ("weight", "bias"):
        m[f"{src}post_extract_proj.{leaf}"] = f"{dst}proj_to_input.layer.{leaf}"
        m[f"{src}encoder.layer_norm.{leaf}"] = f"{dst}encoder.ln.{leaf}"
    for leaf in ("bias", "weight_g", "weight_v"):
        m[f"{src}encoder.pos_conv.0.{leaf}"] = f"{dst}encoder.pos_conv.conv.1.{leaf}"
    if not ctc:
        for name in ("project_q", "final_proj"):
            for leaf in ("weight", "bias"):
                m[f"{name}.{leaf}"] = f"{name}.layer.{leaf}"
        return m
    for i in range(len(CONV_FEATURES[sr])):
        m[f"{src}feature_extractor.conv_layers.{i}.0.weight"] = f"{dst}feature_extractor.conv_layers.{i}.0.weight"
    for leaf in ("weight", "bias"):
        m[f"{src}feature_extractor.conv_layers.0.2.{leaf}"] = f"{dst}feature_extractor.conv_layers.0.2.{leaf}"
        m[f"{src}layer_norm.{leaf}"] = f"{dst}layer_norm.{leaf}"
        m[f"w2v_encoder.proj.{leaf}"] = f"proj.{leaf}"
    m[f"{src}mask_emb"] = f"{dst}mask_emb"
    return m


def convert_keys(num_layers, d, ctc=False, sr=16):
    """rename a fairseq state dict (consumed, like the reference pops from it); a mapped key missing from `d` raises
    KeyError as in the reference (wav2vec2.py:154-168)"""
    out = {}
    for k, v in fairseq_key_map(num_layers, ctc, sr).items():
        out[v] = d.pop(k)
    out.update(d)
    return out


def load_fairseq_bin(w2v, bin_file, ctc=False, sr=16):
    """Reference wav2vec2.py:171-186: load a fairseq `.pt`/`.bin` checkpoint (`{"model": state_dict}`) into a drop-in
    model; returns {'missing': [...], 'unexpected': [...]} from a non-strict load."""
    transformer = w2v.encoder.encoder.transformer if ctc else w2v.encoder.transformer
    d = torch.load(bin_file, map_location="cpu")["model"]
    mapped = convert_keys(len(transformer.encoders), d, ctc, sr)
    res = w2v.load_state_dict(mapped, strict=False)
    return {"missing": list(res.missing_keys), "unexpected": list(res.unexpected_keys)}
