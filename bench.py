#!/usr/bin/env python
"""Benchmark of the wav2vec2-base contrastive pre-training step (BASELINE.json: configs[1], "C2" in SURVEY §8):
fwd+bwd of `loss_function(model, inputs)` on synthetic 16 kHz audio, B=6 x 15 s crops per GPU, dropout 0.1
(the reference's defaults), reported as audio-seconds/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload pretrain|ctc] [--model base|large]

`--workload ctc` times BASELINE configs[2] instead (wav2vec2-base CTC fine-tuning step, char vocab 32, B=8 x 15 s ragged
per GPU, 150-char targets, frozen feature encoder as train.py defaults); `--model large` configs[3].

* our arm: audio8_b200 modules (hand-written sm_100a kernels through the C ABI); N>1 under torchrun with DDP/NCCL.
* `--impl reference`: the reference algorithm on the host CPU cores (oracle port of audio8/wav2vec2.py — the
  reference itself cannot be installed: its `mead-baseline` dependency is absent and there is no network), on a
  bounded sample of the same workload.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLE_RATE = 16000
CROP_S = 15
L = SAMPLE_RATE * CROP_S  # 240000 samples -> 749 frames
B_PER_GPU = 6  # what AudioFileDataset emits for 15 s crops at tokens_per_batch=1.4M (data.py:417-426)
N_VARS, N_NEG = 640, 100
GFLOP_PER_AUDIO_S = 46.0  # algorithmic fwd+bwd, SURVEY §8(d)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p.get("bf16_tflops_sustained", 1400.0), p.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, sustained)"
    except Exception:
        return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region, without perturbing it: NVML queries (in-process
    or through nvidia-smi) take driver-wide locks and were measured to stall concurrent kernel launches by 40-130 ms on
    these boxes, which drained the GPU in the middle of the timed loop.  The samples are therefore taken by the main
    thread right after the LAST step of the region has been enqueued, while the GPU is still executing the queued steps
    (the host runs several steps ahead of the device), i.e. under the region's load but with no launch in flight."""

    def __init__(self, index):
        self.index = index
        self.samples = []  # (sm_mhz, max_mhz, set(reasons))
        self.power = []    # (watts drawn, enforced limit) per sample
        self.how = None
        self.h = None

    def start(self):
        if os.environ.get("A8_NO_CLOCKS"):
            return
        try:
            import pynvml as N
            N.nvmlInit()
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except ValueError:
                    pass
            self.N, self.h = N, N.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self.how = "nvml"
        except Exception:
            self.how = "nvidia-smi"

    def sample(self):
        if self.how == "nvml":
            N = self.N
            names = [("hw_slowdown", N.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", N.nvmlClocksEventReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", N.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", N.nvmlClocksEventReasonSwPowerCap)]
            # every other bit NVML may raise is reported under its own name too: a clock below the maximum should come
            # with its reason (power draw against the enforced limit is sampled for the same purpose)
            for attr, nm in (("nvmlClocksEventReasonHwPowerBrakeSlowdown", "hw_power_brake_slowdown"),
                             ("nvmlClocksEventReasonApplicationsClocksSetting", "applications_clocks_setting"),
                             ("nvmlClocksEventReasonSyncBoost", "sync_boost"),
                             ("nvmlClocksEventReasonDisplayClockSetting", "display_clock_setting")):
                if hasattr(N, attr):
                    names.append((nm, getattr(N, attr)))
            try:
                sm = float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM))
                bits = int(N.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.samples.append((sm, self.mx, {n for n, b in names if bits & b}))
                try:
                    self.power.append((N.nvmlDeviceGetPowerUsage(self.h) / 1e3, N.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1e3))
                except Exception:
                    pass
            except Exception:
                pass
        elif self.how == "nvidia-smi":
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
                f = [x.strip() for x in out.split(",")]
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                self.samples.append((float(f[0]), float(f[1]), {n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")}))
            except Exception:
                pass

    def sample_while_busy(self, done_event, max_samples=4):
        """a few samples while the stream is still working through the queued steps"""
        n = 0
        while n < max_samples and not done_event.query():
            self.sample()
            n += 1
            time.sleep(0.002)
        self.in_region = n

    def stop(self):
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        reasons = set()
        for x in self.samples:
            reasons |= x[2]
        sm = [x[0] for x in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.samples[0][1] if self.samples else None,
                "samples": len(self.samples), "samples_while_gpu_busy": getattr(self, "in_region", 0),
                "reasons": sorted(reasons), "source": self.how,
                "power_w": float(np.median([p[0] for p in self.power])) if self.power else None,
                "power_limit_w": self.power[0][1] if self.power else None,
                "when": "after the last timed step was enqueued, while the GPU was still executing the timed region"}


class GemmProfiler:
    """CUDA events around every a8_gemm launch (on the launching stream)"""

    def __init__(self):
        self.pairs = []
        self.cur = None

    def begin(self, kind, flops):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.cur = (e0, e1, flops)

    def end(self):
        e0, e1, flops = self.cur
        e1.record()
        self.pairs.append((e0, e1, flops))

    def summary(self):
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b, _ in self.pairs)
        return ms, sum(f for _, _, f in self.pairs), len(self.pairs)


# ------------------------------------------------------------------------------------------------------------------
B_CTC, S_CTC, V_CTC = 8, 150, 32  # BASELINE configs[2] per GPU


def workload_name(args):
    if args.workload == "ctc":
        return "wav2vec2-base (12L d=768) CTC fine-tune fwd+bwd, char vocab 32, 150-char targets, ragged 15 s crops, frozen feature encoder, dropout 0.1"
    m = "wav2vec2-large (24L d=1024)" if args.model == "large" else "wav2vec2-base (12L d=768)"
    return m + " contrastive pretrain fwd+bwd, G=2 V=320 K=100, dropout 0.1"


def metric_name(args):
    if args.workload == "ctc":
        return "wav2vec2-base CTC fine-tune audio-sec/sec fwd+bwd"
    return ("wav2vec2-large" if args.model == "large" else "wav2vec2-base") + " pretrain audio-sec/sec fwd+bwd"


def bench_config(args, world, extra=None):
    """the `config` object shared by both arms (same keys, same workload string)"""
    B = B_CTC if args.workload == "ctc" else B_PER_GPU
    cfg = {"workload": workload_name(args), "batch_per_gpu": B, "crop_s": CROP_S, "global_batch": world * B,
           "parallelism": f"dp{world}"}
    if extra:
        cfg.update(extra)
    return cfg


def ctc_batch(B, gen):
    """synthetic fine-tuning batch: ragged utterances (70-100 % of 15 s, one full-length), 150-char targets"""
    x = torch.randn(B, L, generator=gen) * 0.1
    in_len = torch.randint(int(0.7 * L), L + 1, (B,), generator=gen)
    in_len[0] = L
    pad_mask = torch.arange(L)[None, :] < in_len[:, None]
    x = x * pad_mask
    targets = torch.randint(4, V_CTC, (B, S_CTC), generator=gen)
    tl = torch.full((B,), S_CTC, dtype=torch.long)
    return x, pad_mask, targets, tl


def reference_step_fn(args, device="cpu", threads=None, dtype_mode="fp32"):
    """The reference algorithm (oracle port of audio8/wav2vec2.py:377-392,927-952 / :696-770 + ctc.py:186-206), fwd+bwd,
    dropout 0.1 as in the reference's defaults.  device="cpu": the CPU arm; device=cuda: the eager-PyTorch GPU incumbent
    (dtype_mode fp32 | tf32 | bf16-autocast)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_ctc
    import ref_params as P
    import ref_wav2vec2 as R
    if threads:
        torch.set_num_threads(threads)
    large = args.model == "large"
    kw = dict(d_model=1024, num_layers=24, d_ff=4096) if large else {}
    heads, layers = (16, 24) if large else (12, 12)
    if args.workload == "ctc":
        sd = P.acoustic_state_dict(V_CTC, seed=0, **kw)
    else:
        sd = P.pretrain_state_dict(seed=0, **kw)
    sd = {k: v.to(device).requires_grad_(True) for k, v in sd.items()}
    g = torch.Generator().manual_seed(0)
    if device != "cpu":
        torch.backends.cuda.matmul.allow_tf32 = dtype_mode == "tf32"
        torch.backends.cudnn.allow_tf32 = dtype_mode == "tf32"
    import contextlib
    cast = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if dtype_mode == "bf16" else contextlib.nullcontext

    def step_pretrain(batch):
        x = (torch.randn(batch, L, generator=g) * 0.1).to(device)
        T = R.conv_out_lengths(L, R.CONV_FEATURES[16])[-1]
        tmask = R.create_mask((batch, T), 0.65, 10)
        Tm = int(tmask[0].sum())
        for _ in range(layers):
            np.random.random()
        idx = R.sample_negative_indices(batch, Tm, N_NEG)
        noise = -torch.empty(batch * Tm * 2, 320, device=device).exponential_().log()
        with cast():
            st = R.pretrain_loss(sd, x, tmask, idx, n_vars=N_VARS, gumbel_noise=noise, num_heads=heads, num_layers=layers,
                                 dropout=0.1, dropout_input=0.1, dropout_features=0.1)
        st["loss"].backward()
        for v in sd.values():
            v.grad = None
        return st["loss"]

    def step_ctc(batch):
        x, pad_mask, targets, tl = ctc_batch(batch, g)
        x, pad_mask = x.to(device), pad_mask.to(device)
        T = R.conv_out_lengths(L, R.CONV_FEATURES[16])[-1]
        tm = R.create_mask((batch, T), 0.5, 10)
        cm = R.create_mask((batch, 1024 if large else 768), 0.1, 64)
        for _ in range(layers):
            np.random.random()
        with cast():
            lp, fmask = R.acoustic_forward(sd, x, pad_mask, heads, layers, tm, cm, dropout=0.1, freeze_fx=True)
        loss = ref_ctc.ctc_loss_reference(lp.float().transpose(1, 0), fmask.sum(-1).cpu(), targets.to(device), tl, 0, 1, 2)
        loss.backward()
        for v in sd.values():
            v.grad = None
        return loss

    return step_ctc if args.workload == "ctc" else step_pretrain


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step = reference_step_fn(args, "cpu", threads)
    batch = B_CTC if args.workload == "ctc" else B_PER_GPU  # the arm's own per-GPU batch: same config as ours
    # bounded sample: a probe step of one utterance sizes the batch so that the whole run stays within ~4 minutes
    step(1)
    t0 = time.perf_counter()
    step(1)
    t1 = time.perf_counter() - t0
    batch = int(max(1, min(batch, 240.0 / (max(args.steps + args.warmup, 1) * t1))))
    for _ in range(args.warmup):
        step(batch)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(batch).item()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    v = batch * CROP_S / dt
    sample = (f"{args.steps} steps of B={batch} x {CROP_S} s (one GPU's batch) on the host CPU, {threads} threads "
              "(oracle port of the reference algorithm, fp32, dropout 0.1)")
    out = {
        "impl": "reference", "metric": metric_name(args), "value": v, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, max(args.gpus, 1)),
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def gpu_incumbent(args, dev, B):
    """SURVEY §8d's 'honest GPU incumbent': the reference algorithm (oracle port, plain eager PyTorch: cuDNN convs, cuBLAS
    matmuls, ATen softmax/LayerNorm/CTC) on the same B200, same batch, CUDA events; informational."""
    out = {}
    for mode in ("fp32", "tf32", "bf16"):
        try:
            step = reference_step_fn(args, str(dev), None, mode)
            for _ in range(2):
                step(B)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5
            e0.record()
            for _ in range(n):
                step(B)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out[mode] = {"ms_per_step": ms, "value": B * CROP_S / (ms * 1e-3), "unit": "audio-s/s"}
        except Exception as e:  # informational leg: never takes the bench line down
            out[mode] = {"error": repr(e)[:200]}
        finally:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = True
            torch.cuda.empty_cache()
    out["what"] = ("oracle port of the reference modules under eager PyTorch on this GPU (fp32 with TF32 off / TF32 on / "
                   f"bf16 autocast), B={B} x {CROP_S} s, dropout 0.1, 5 steps after 2 warm-up, CUDA events; host draws included")
    return out


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    # stdout carries exactly ONE line, the JSON: anything libraries print on file descriptor 1 meanwhile (NCCL's version
    # banner, for one) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(line, flush=True)


def _run_ours(args):
    import torch.distributed as dist
    from audio8_b200 import _lib, ops
    from audio8_b200 import wav2vec2 as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    force_dp = os.environ.get("A8_DP_FORCE") == "1"  # diagnostics: the data-parallel wrapper (arena, hooks) on one GPU
    if force_dp and world == 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    if world > 1:
        import datetime
        # a short collective timeout: a rank that falls out of step must fail in minutes, not hold the box for ten
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    torch.manual_seed(1234 + rank)
    np.random.seed(1234 + rank)

    os.environ.setdefault("A8_GRAPH_STRICT", "1")  # a failed CUDA-graph capture must fail the bench, not slow it down
    ctc = args.workload == "ctc"
    large = args.model == "large"
    mkw = dict(d_model=1024, num_heads=16, num_layers=24, d_ff=4096) if large else {}
    # algorithmic fwd+bwd GFLOP per audio-second; CTC fine-tuning: the frozen conv stack counts once
    gflop_per_audio_s = (3 * 39.8 if large else GFLOP_PER_AUDIO_S) if not ctc else (46.0 - 2 * 0.32 * 15.3)
    if ctc:  # BASELINE configs[2]: create_acoustic_model(32) as train.py builds it (freeze_fx=True), encoder trainable
        from audio8_b200.ctc import CTCLoss, Offsets
        Offsets.GO, Offsets.PAD = 0, 1  # train.py:22-27
        model = W.create_acoustic_model(V_CTC, **mkw).to(dev)
        model.freeze = False
        crit = CTCLoss()
    else:  # the headline: wav2vec2-base defaults, 12L d=768, dropout 0.1, G=2 V=320 (BASELINE configs[1]; large = configs[3])
        model = W.create_model(**mkw).to(dev)
        # span mask + negatives: drawn on the device inside the step's CUDA graphs (csrc/draws.cu; the reference's
        # distributions, seeded by torch's CUDA generator), or --draws host: numpy's global generator in the reference's call
        # order (its exact numbers), drawn one step ahead on a helper thread
        W.set_device_draws(args.draws == "device")
        W.set_prefetch_draws(True)
    model.train()
    loss_fn = W.create_loss(N_VARS, N_NEG)
    net = model
    dp_name = "single process"
    if world > 1 or force_dp:
        if os.environ.get("A8_DP", "arena") == "ddp":
            # stock DistributedDataParallel as in pretrain.py:158; gradients as views of few large buckets
            net = torch.nn.parallel.DistributedDataParallel(
                model, device_ids=[local], output_device=local,
                gradient_as_bucket_view=os.environ.get("A8_DDP_BUCKET_VIEW", "1") != "0",
                bucket_cap_mb=int(os.environ.get("A8_DDP_BUCKET_MB", "128")))
            dp_name = "torch DistributedDataParallel (bucket views, 128 MB buckets)"
        else:
            # the package's data-parallel wrapper: gradients are written into one contiguous arena and all-reduced in
            # place under the rest of backward (audio8_b200/parallel.py); same interface as DDP
            from audio8_b200.parallel import DataParallel
            net = DataParallel(model, force_sync=force_dp)
            dp_name = "audio8_b200.parallel.DataParallel (gradient arena, in-place all-reduce: NVSwitch multicast kernel, else NCCL)"
    B = B_CTC if ctc else B_PER_GPU
    lib = _lib.load()
    gen = torch.Generator().manual_seed(1234 + rank)
    params = [p for p in model.parameters()]
    if ctc:
        xh, pmh, tgh, tlh = ctc_batch(B, gen)
        host_in = tuple(t.pin_memory() for t in (xh, pmh, tgh))
        dev_in = tuple(t.to(dev) for t in host_in)
        h2d_bytes = sum(t.numel() * t.element_size() for t in host_in)

        def step(inp, module=None):
            x, pm, tg = inp
            lp, fmask = (net if module is None else module)(x, pm)
            loss = crit(lp.transpose(1, 0), fmask.sum(-1), tg, tlh)
            loss.backward()
            for p in params:
                p.grad = None
            return loss
    else:
        dev_in = (torch.randn(B, L, device=dev) * 0.1,)
        host_in = ((torch.randn(B, L, generator=gen) * 0.1).pin_memory(),)
        h2d_bytes = B * L * 4

        def step(inp, module=None):
            loss = loss_fn(net if module is None else module, inp[0])
            loss.backward()
            for p in params:
                p.grad = None
            return loss

    def host_batches():  # the data loader of this benchmark: the same pinned host batch, forever
        while True:
            yield host_in

    from audio8_b200.feed import DeviceFeed
    feed = DeviceFeed(host_batches(), dev, depth=2)  # public input feed: H2D on a copy stream, one batch ahead

    def upload():
        return next(feed)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # set-up steps, NOT the warm-up the caller asked for: CUDA-graph capture happens on the 2nd step and torch's caching
    # allocator needs ~10 steps of varying masked-row counts before it stops calling cudaMalloc; reported as extra_warmup
    extra_warm = 12
    for _ in range(extra_warm):
        step(dev_in)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):  # the W untimed warm-up steps of the contract
        step(dev_in)
    # Python's cyclic GC: a full (generation-2) collection walks every live object of the process (~100 ms with torch
    # loaded) and lands inside whichever step happens to trip its allocation counter.  Like Megatron-style trainers do,
    # freeze what exists after warm-up and collect by hand between the timed regions, not inside them.
    import gc
    gc.collect()
    gc.freeze()  # the young generations stay enabled: they are cheap and reclaim whatever cycles a step leaves behind
    if args.ncu_step:
        # `ncu --profile-from-start off ... bench.py --ncu-step`: exactly ONE warmed-up step inside the profiler range
        # (numbers printed by a run under ncu are never bench values: nothing is printed)
        eager = os.environ.get("A8_NCU_EAGER")
        if eager:
            from audio8_b200 import graphs
            graphs.set_enabled(False)
            step(dev_in)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(dev_in)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # ---- device-resident timing: exactly K steps between barriers, CUDA events, max over ranks
    barrier()
    n0 = lib.a8_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_t, step_ev = [], []
    e0.record()
    for _ in range(args.steps):
        th = time.perf_counter()
        step(dev_in)
        host_t.append((time.perf_counter() - th) * 1e3)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        step_ev.append(ev)
    e1.record()
    if rank == 0:
        clocks.sample_while_busy(e1)
    barrier()
    launches = lib.a8_launch_count() - n0
    if rank == 0:
        sys.stderr.write("host enqueue ms per step: " + " ".join(f"{t:.1f}" for t in host_t) + "\n")
        prev, gpu_t = e0, []
        for ev in step_ev:
            gpu_t.append(prev.elapsed_time(ev))
            prev = ev
        sys.stderr.write("gpu ms per step (event to event): " + " ".join(f"{t:.1f}" for t in gpu_t) + "\n")
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    clk = clocks.stop() if rank == 0 else None
    value = world * B * CROP_S / (ms_per_step * 1e-3)

    # ---- end to end through the public API: pinned host input -> device every step, loss read back every step
    gc.collect()
    barrier()
    for _ in range(8):  # this loop's own warm-up: the per-step input tensor changes the allocator's request sequence
        loss = step(upload())  # same statement shape as the timed loop (object lifetimes)
        loss_val = loss.item()
    gc.collect()
    barrier()
    t0 = time.perf_counter()
    e2e_t = []
    for _ in range(args.steps):
        ts = time.perf_counter()
        loss = step(upload())
        loss_val = loss.item()
        e2e_t.append((time.perf_counter() - ts) * 1e3)
    if rank == 0:
        sys.stderr.write("e2e ms per step (host clock, loss.item() each step): " + " ".join(f"{t:.1f}" for t in e2e_t) + "\n")
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = world * B * CROP_S / (dt.item() / args.steps)

    # ---- the same end-to-end loop with the OTHER draw mode's host path (informational, one GPU only): numpy draws in the
    # reference's order (prefetched) + the padded index uploads, i.e. the mode whose masks / negatives are bit-identical to
    # the reference's
    e2e_host = None
    if not ctc and world == 1 and args.draws == "device" and not force_dp:
        W.set_device_draws(False)
        try:
            for _ in range(10):  # eager step, capture of the host-draw segments, replays
                loss = step(upload())
                loss.item()
            gc.collect()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                loss = step(upload())
                loss.item()
            e2e_host = {"value": B * CROP_S / ((time.perf_counter() - t0) / args.steps), "unit": "audio-s/s",
                        "what": "the e2e loop with --draws host: numpy global-RNG draws in the reference's call order "
                                "(bit-identical masks / negatives), prefetched one step ahead, index lists uploaded"}
        finally:
            W.set_device_draws(True)
            W._HostDraws._pending = None
        for _ in range(2):
            step(dev_in)

    # ---- host-side enqueue time of one step (queue empty at start, no sync inside): the launch-overhead floor
    gc.collect()
    barrier()
    t0 = time.perf_counter()
    step(dev_in)
    host_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    # ---- the whole training step as pretrain.py:176-184 runs it: fwd + bwd + clip_grad_norm_(1.0) + AdamW + zero_grad,
    # with the optimizer side done by audio8_b200.optim.FusedAdamW (two multi-tensor launches); informational key
    from audio8_b200.optim import FusedAdamW
    opt = FusedAdamW(model.parameters(), lr=2.0e-4, weight_decay=1.0e-2)

    def train_step():
        loss = loss_fn(net, dev_in[0]) if not ctc else None
        if ctc:
            lp, fmask = net(dev_in[0], dev_in[1])
            loss = crit(lp.transpose(1, 0), fmask.sum(-1), dev_in[2], tlh)
        loss.backward()
        opt.step(clip=1.0)
        for p in params:
            p.grad = None

    for _ in range(3):
        train_step()
    gc.collect()
    barrier()
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    o0.record()
    for _ in range(args.steps):
        train_step()
    o1.record()
    barrier()
    oms = torch.tensor([o0.elapsed_time(o1)], device=dev)
    if world > 1:
        dist.all_reduce(oms, op=dist.ReduceOp.MAX)
    opt_ms = oms.item() / args.steps
    with_opt = {"ms_per_step": opt_ms, "value": world * B * CROP_S / (opt_ms * 1e-3), "unit": "audio-s/s",
                "optimizer_ms_per_step": opt_ms - ms_per_step,
                "what": "fwd + bwd + fused clip_grad_norm_(1.0) + AdamW (audio8_b200.optim.FusedAdamW) + zero_grad, "
                        "device-resident inputs, CUDA events"}
    del opt
    for p in params:
        p.grad = None
    torch.cuda.empty_cache()

    out = None
    result_line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (tcgen05 GEMM): CUDA events around every launch of 2 further steps
        from audio8_b200 import graphs
        graphs.set_enabled(False)  # single launches cannot be bracketed inside a graph replay: eager for these 2 steps
        # rank 0 only: these steps run on the bare module (a DDP-wrapped step here would wait for the other ranks'
        # all-reduce forever)
        step(dev_in, model)
        step(dev_in, model)  # the grouped launch's prepared descriptors are cached on addresses: let the allocator settle
        prof = GemmProfiler()
        ops.backend().profiler = prof
        for _ in range(2):
            # keep the stream busy while the host enqueues the eager step: with an empty queue every (event, launch, event)
            # triple would also time the host's latency between the two driver calls (~5-10 us per launch from Python)
            torch.cuda._sleep(int(40e6))
            step(dev_in, model)
        gemm_ms, gemm_flops, n_gemm = prof.summary()
        ops.backend().profiler = None
        graphs.set_enabled(True)
        tf_peak, hbm_peak, which = peaks()
        traffic, traffic_src = None, None  # DRAM bytes per GEMM launch from the committed ncu capture (profiles/)
        for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as f:
                    tr = json.load(f)
                traffic = tr["gemm_dram_bytes_per_launch"]  # bytes per launch, like `achieved` is per launch
                traffic_src = tr["source"]
                break
            except Exception:
                pass
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        # ---- CPU baseline (oracle port) on a bounded sample, rank 0, N=1 only
        cpu = incumbent = None
        if world == 1 and not args.no_incumbent:
            incumbent = gpu_incumbent(args, dev, B)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cstep = reference_step_fn(args, "cpu", threads)
            cb = 2
            cstep(cb)  # 1 warm-up, then 3 timed steps
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                cstep(cb).item()
            cdt = (time.perf_counter() - t0) / reps
            cpu = {"value": cb * CROP_S / cdt, "unit": "audio-s/s", "cores": threads, "kind": "port",
                   "sample": f"{reps} steps of B={cb} x {CROP_S} s after 1 warm-up (oracle port of the reference algorithm, "
                             "fp32, dropout 0.1)"}
        out = {
            "metric": metric_name(args), "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "extra_warmup": extra_warm,
            "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": bench_config(args, world, {
                "data_parallel": dp_name if getattr(net, "_arena", None) is None else (
                    "audio8_b200.parallel.DataParallel (gradient arena, in-place all-reduce: "
                    + ("NVSwitch multicast kernel a8_allreduce_mc)" if net._arena.switch is not None else "NCCL)")),
                "l2": "inputs+activations per step (>1 GB) exceed the 126 MB L2; no explicit flush",
                "launch": "the step is 4 CUDA-graph segments (front+mask, quantizer branch, encoder, loss), fwd and bwd; "
                          "masked-row lists padded to their worst-case length; "
                          + ("span mask and negatives drawn on the device inside the segments (csrc/draws.cu, --draws device: "
                             "the reference's distributions, not its numpy numbers; e2e_host_draws = the numpy mode)"
                             if (not ctc and W._DEVICE_DRAWS[0]) else
                             "host draws (numpy, the reference's numbers) prefetched one step ahead"),
                "operands": "bf16 / packed operand copies of the parameters are rebuilt when a parameter changes (version "
                            "counter), i.e. once per optimizer step; the fwd+bwd loop of `value` never changes them, "
                            "step_with_optimizer rebuilds them every step",
                "gc": "gc.freeze() after warm-up (full collections no longer walk the long-lived heap)"}),
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "last_loss": loss_val},
            "gpu_launches": int(launches),
            "gpu_launches_per_step": launches / args.steps,
            "host_enqueue_ms_per_step": host_ms,
            "e2e_host_draws": e2e_host,
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05)", "achieved": achieved, "peak": tf_peak,
                         "unit": "TFLOP/s", "frac": achieved / tf_peak, "traffic": traffic, "traffic_unit": "bytes per launch",
                         "traffic_source": traffic_src, "peak_source": which,
                         "launches_per_step": n_gemm / 2, "gemm_ms_per_step": gemm_ms / 2,
                         "gemm_share_of_step": (gemm_ms / 2) / ms_per_step,
                         "algorithmic_gflop_per_step": gemm_flops / 2 / 1e9,
                         "measured_on": "2 extra eager steps after the timed region, CUDA events around each launch, the "
                                        "stream kept busy ahead of the host (a 20 ms spin kernel in front of each step)",
                         "model_frac_of_tensor_roofline": value / world * gflop_per_audio_s / 1e3 / tf_peak},
            "cpu_baseline": cpu,
            "gpu_incumbent": incumbent,
            "step_with_optimizer": with_opt,
        }
        result_line = json.dumps(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result_line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--draws", default="device", choices=["device", "host"],
                    help="pretrain workload: where the span mask and the negatives are drawn (see audio8_b200.wav2vec2.set_device_draws)")
    ap.add_argument("--no-incumbent", action="store_true", help="skip the eager-PyTorch GPU incumbent leg")
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "ctc"],
                    help="pretrain = BASELINE configs[1] (the driver's line); ctc = configs[2] (CTC fine-tuning step)")
    ap.add_argument("--model", default="base", choices=["base", "large"],
                    help="base = the headline workload (BASELINE configs[1]); large = configs[3] (not the driver's line)")
    ap.add_argument("--ncu-step", action="store_true", help="profile exactly one step (cudaProfilerStart/Stop) and exit")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
