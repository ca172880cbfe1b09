"""Data-parallel path: one process per rank, batch sharded by utterance, gradients averaged by
DistributedDataParallel — here world_size 2 over `gloo` on CPU with the ABI emulation standing in for the kernels
(the host-side orchestration, the autograd Functions and their interaction with DDP's reducer hooks are what is
being tested; NCCL replaces gloo on the GPU box, `bench.py --gpus N`).

Reference behaviour (pretrain.py:158,178-179): `DistributedDataParallel(model)`, per-rank `loss_function(model, x)`,
`.backward()` — DDP's plain average of per-rank gradients; masks / negatives / perplexity are per-rank quantities.
"""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

TINY = dict(d_model=128, num_heads=2, num_layers=1, d_ff=256, final_dim=64, num_vq_vars=24, num_vq_groups=2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _local_grads(model, loss_fn, x, seed):
    np.random.seed(seed)
    torch.manual_seed(seed)
    model.zero_grad(set_to_none=True)
    loss = loss_fn(model, x)
    loss.backward()
    return loss.item(), {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


def _worker(rank, world, port, out_dir, no_sync):
    for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu
        from audio8_b200 import ops
        from audio8_b200 import wav2vec2 as W
        ops.set_backend(emu.EmuOps())
        torch.manual_seed(0)  # identical initial weights on every rank (DDP would broadcast rank 0's anyway)
        model = W.create_model(dropout=0.0, dropout_input=0.0, dropout_features=0.0, **TINY).train()
        loss_fn = W.create_loss(TINY["num_vq_vars"] * TINY["num_vq_groups"], 10)
        xs = [torch.randn(2, 5000, generator=torch.Generator().manual_seed(100 + r)) * 0.1 for r in range(world)]
        # expected: average over ranks of the per-rank gradients, computed locally without DDP
        want, losses = None, []
        for r in range(world):
            l, g = _local_grads(model, loss_fn, xs[r], 7 + r)
            losses.append(l)
            want = g if want is None else {k: want[k] + g[k] for k in g}
        want = {k: v / world for k, v in want.items()}
        ddp = torch.nn.parallel.DistributedDataParallel(model)
        if no_sync:
            # gradient accumulation (train.py:301): micro-step under no_sync() must not all-reduce
            with ddp.no_sync():
                l0, g_local = _local_grads(ddp, loss_fn, xs[rank], 7 + rank)
            g_local = {k.replace("module.", "", 1): v for k, v in g_local.items()}
            solo = _local_grads(model, loss_fn, xs[rank], 7 + rank)[1]
            for k in solo:
                assert torch.allclose(g_local[k], solo[k], rtol=1e-5, atol=1e-7), f"no_sync changed grad {k}"
        l, got = _local_grads(ddp, loss_fn, xs[rank], 7 + rank)
        assert abs(l - losses[rank]) <= 1e-5 * abs(l), (l, losses[rank])
        got = {k.replace("module.", "", 1): v for k, v in got.items()}
        assert set(got) == set(want)
        for k in want:
            err = (got[k] - want[k]).abs().max().item()
            scale = want[k].abs().max().item() + 1e-8
            assert err <= 1e-4 * scale + 1e-7, f"rank {rank} grad {k}: {err:.3g} vs scale {scale:.3g}"
        # every rank holds the same averaged gradient
        flat = torch.cat([got[k].reshape(-1) for k in sorted(got)])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        for g in gathered:
            assert torch.equal(g, flat)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_ddp_gloo_world2_matches_mean_of_rank_gradients(tmp_path, no_sync=True):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), no_sync), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def _ctc_worker(rank, world, port, out_dir):
    """fine-tuning normalisation (train.py:318-323): DDP mean x num_gpus / global target count"""
    for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu
        from audio8_b200 import ops
        from audio8_b200 import wav2vec2 as W
        from audio8_b200.ctc import CTCLoss
        ops.set_backend(emu.EmuOps())
        torch.manual_seed(0)
        model = W.create_acoustic_model(32, d_model=128, num_heads=2, num_layers=1, d_ff=256, dropout=0.0,
                                        timestep_masking=0.0, channel_masking=0.0).train()
        model.freeze = False
        crit = CTCLoss()
        B = 2 + rank  # unequal per-rank batches
        g = torch.Generator().manual_seed(50 + rank)
        x = torch.randn(B, 8000, generator=g) * 0.1
        pad_mask = torch.ones(B, 8000, dtype=torch.bool)
        targets = torch.randint(4, 32, (B, 5), generator=g)
        tl = torch.full((B,), 5, dtype=torch.int64)
        # train.py:266-268 wraps with find_unused_parameters=True (frozen feature encoder, unused mask_emb)
        ddp = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True)
        lp, fmask = ddp(x, pad_mask)
        loss = crit(lp.transpose(1, 0), fmask.sum(-1), targets, tl)
        loss.backward()
        n = torch.tensor([float(tl.sum())])
        dist.all_reduce(n)
        scale = world / n.item()
        grads = torch.cat([p.grad.reshape(-1) * scale for p in model.parameters() if p.grad is not None])
        assert torch.isfinite(grads).all() and grads.abs().max() > 0
        gathered = [torch.empty_like(grads) for _ in range(world)]
        dist.all_gather(gathered, grads)
        assert torch.equal(gathered[0], gathered[1])
        # the package's own wrapper on the same step: same averaged gradients as DistributedDataParallel
        from audio8_b200.parallel import DataParallel
        del ddp
        model.zero_grad(set_to_none=True)
        net = DataParallel(model)
        lp, fmask = net(x, pad_mask)
        crit(lp.transpose(1, 0), fmask.sum(-1), targets, tl).backward()
        grads2 = torch.cat([p.grad.reshape(-1) * scale for p in model.parameters() if p.grad is not None])
        assert grads2.shape == grads.shape
        assert (grads2 - grads).abs().max().item() <= 1e-4 * grads.abs().max().item() + 1e-7
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_ddp_gloo_world2_ctc_unequal_batches(tmp_path):
    world = 2
    mp.spawn(_ctc_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def _arena_worker(rank, world, port, out_dir):
    """audio8_b200.parallel.DataParallel (gradient arena + one in-place all-reduce) against the mean of the per-rank
    gradients, incl. a no_sync() micro-step followed by a synchronised one (gradient accumulation)"""
    for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import emu
        from audio8_b200 import ops
        from audio8_b200 import wav2vec2 as W
        from audio8_b200.parallel import DataParallel
        ops.set_backend(emu.EmuOps())
        torch.manual_seed(rank)  # different initial weights per rank: the wrapper must broadcast rank 0's
        model = W.create_model(dropout=0.0, dropout_input=0.0, dropout_features=0.0, **TINY).train()
        loss_fn = W.create_loss(TINY["num_vq_vars"] * TINY["num_vq_groups"], 10)
        net = DataParallel(model)
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert torch.equal(gathered[0], gathered[1]), "parameters were not broadcast"
        xs = [torch.randn(2, 5000, generator=torch.Generator().manual_seed(100 + r)) * 0.1 for r in range(world)]
        want = None
        for r in range(world):
            _, g = _local_grads(model, loss_fn, xs[r], 7 + r)
            want = g if want is None else {k: want[k] + g[k] for k in g}
        want = {k: v / world for k, v in want.items()}
        for rep in range(2):  # twice: the second step reuses the arena blocks
            _, got = _local_grads(net, loss_fn, xs[rank], 7 + rank)
            got = {k.replace("module.", "", 1): v for k, v in got.items()}
            assert set(got) == set(want)
            for k in want:
                err = (got[k] - want[k]).abs().max().item()
                scale = want[k].abs().max().item() + 1e-8
                assert err <= 1e-4 * scale + 1e-7, f"rank {rank} step {rep} grad {k}: {err:.3g} vs scale {scale:.3g}"
        assert net._arena is not None and net._arena.used > 0, "the transformer layers did not use the gradient arena"
        in_arena = sum(1 for p in model.parameters() if p.grad is not None and
                       net._arena.buf.data_ptr() <= p.grad.data_ptr() < net._arena.buf.data_ptr() + 4 * net._arena.buf.numel())
        n_grads = sum(1 for p in model.parameters() if p.grad is not None)
        assert in_arena == n_grads, f"only {in_arena} of {n_grads} gradients alias the arena (all of them should)"
        # accumulation: micro-step under no_sync() + synchronised micro-step = mean over ranks of the SUM of both
        model.zero_grad(set_to_none=True)
        np.random.seed(7 + rank)
        torch.manual_seed(7 + rank)
        with net.no_sync():
            loss_fn(net, xs[rank]).backward()
        np.random.seed(7 + rank)
        torch.manual_seed(7 + rank)
        loss_fn(net, xs[rank]).backward()
        for k, p_ in model.named_parameters():
            if p_.grad is None:
                continue
            err = (p_.grad - 2 * want[k]).abs().max().item()
            scale = want[k].abs().max().item() + 1e-8
            assert err <= 2e-4 * scale + 1e-7, f"accumulated grad {k}: {err:.3g} vs scale {scale:.3g}"
        # device-side draws under the wrapper: the span mask / negatives a rank draws follow its torch seed and
        # the call-site counter, so the local runs reproduce them; expected = mean over ranks as above
        from audio8_b200 import functional as Fn
        W.set_device_draws(True)
        try:
            want = None
            for r in range(world):
                Fn._site[0] = 99
                _, g = _local_grads(model, loss_fn, xs[r], 7 + r)
                want = g if want is None else {k: want[k] + g[k] for k in g}
            want = {k: v / world for k, v in want.items()}
            Fn._site[0] = 99
            _, got = _local_grads(net, loss_fn, xs[rank], 7 + rank)
            got = {k.replace("module.", "", 1): v for k, v in got.items()}
            assert set(got) == set(want)
            for k in want:
                err = (got[k] - want[k]).abs().max().item()
                scale = want[k].abs().max().item() + 1e-8
                assert err <= 1e-4 * scale + 1e-7, f"device draws, rank {rank} grad {k}: {err:.3g} vs scale {scale:.3g}"
        finally:
            W.set_device_draws(False)
        _layerdrop_body(rank, world, torch.device("cpu"), False)  # LayerDrop under the wrapper (defined below)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_arena_data_parallel_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_arena_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def _layerdrop_worker(rank, world, port, out_dir, backend):
    """with LayerDrop every rank drops DIFFERENT layers (per-rank numpy RNG).  The arena layout is planned from
    the module structure, so a layer's block sits at the same offset on every rank whatever order backward reaches it
    in; a rank that skipped a layer contributes zeros.  Expected: mean over ranks of the per-rank gradients (None = 0)."""
    for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    cuda = backend == "nccl"
    if cuda:
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        torch.set_num_threads(2)
        dev = torch.device("cpu")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _layerdrop_body(rank, world, dev, cuda)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def _layerdrop_body(rank, world, dev, cuda):
    if True:
        from audio8_b200 import ops
        from audio8_b200 import wav2vec2 as W
        from audio8_b200.parallel import DataParallel
        if not cuda:
            import emu
            ops.set_backend(emu.EmuOps())
        cfg = dict(TINY, num_layers=4)
        torch.manual_seed(0)
        model = W.create_model(dropout=0.0, dropout_input=0.0, dropout_features=0.0, layer_drop=0.5, **cfg).to(dev).train()
        loss_fn = W.create_loss(cfg["num_vq_vars"] * cfg["num_vq_groups"], 10)
        L = 16000 if cuda else 5000
        xs = [(torch.randn(2, L, generator=torch.Generator().manual_seed(100 + r)) * 0.1).to(dev) for r in range(world)]
        import ref_wav2vec2 as R
        T = R.conv_out_lengths(L, R.CONV_FEATURES[16])[-1]
        seeds, seen = [], set()
        for sd_ in range(1, 200):  # two numpy seeds whose LayerDrop outcomes differ (and drop some, not all, layers)
            np.random.seed(sd_)
            R.create_mask((2, T), 0.65, 10)
            act = tuple(np.random.random() >= 0.5 for _ in range(cfg["num_layers"]))
            if 0 < sum(act) < len(act) and act not in seen:
                seen.add(act)
                seeds.append(sd_)
            if len(seeds) == world:
                break
        want, dropped = None, []
        for r in range(world):
            _, g = _local_grads(model, loss_fn, xs[r], seeds[r])
            dropped.append(sorted(k for k, p_ in model.named_parameters() if k not in g))
            g = {k: g.get(k, torch.zeros_like(p_)) for k, p_ in model.named_parameters()}
            want = g if want is None else {k: want[k] + g[k] for k in g}
        assert dropped[0] != dropped[1], "pick seeds that drop different layers on the two ranks"
        want = {k: v / world for k, v in want.items()}
        net = DataParallel(model)
        tol = 2e-3 if cuda else 1e-4  # CUDA: fp32 atomics order + bf16 operand rounding is deterministic; atomics are not
        for rep in range(3 if cuda else 2):  # (CPU: the emulated GEMMs make a step cost ~10 s per rank)
            _, got = _local_grads(net, loss_fn, xs[rank], seeds[rank])
            got = {k.replace("module.", "", 1): v for k, v in got.items()}
            assert set(got) == set(want), "a rank that dropped a layer must still hold that layer's averaged gradient"
            for k in want:
                err = (got[k] - want[k]).abs().max().item()
                scale = want[k].abs().max().item() + 1e-8
                assert err <= tol * scale + 1e-7, f"rank {rank} step {rep} grad {k}: {err:.3g} vs scale {scale:.3g}"


def _nccl_worker(rank, world, port, out_dir, exchange="nccl"):
    """the CUDA kernels + NCCL: audio8_b200.parallel.DataParallel against the mean of the per-rank gradients computed
    without any wrapper, over 4 steps (eager, CUDA-graph capture, replays)"""
    for p in (ROOT, HERE, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["A8_ALLREDUCE"] = exchange  # "switch": the multicast kernel even at world size 2 (NCCL when unavailable)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from audio8_b200 import wav2vec2 as W
        from audio8_b200.parallel import DataParallel
        cfg = dict(d_model=256, num_heads=4, num_layers=3, d_ff=1024, final_dim=128, num_vq_vars=64, num_vq_groups=2)
        torch.manual_seed(0)
        model = W.create_model(dropout=0.0, dropout_input=0.0, dropout_features=0.0, **cfg).to(dev).train()
        loss_fn = W.create_loss(128, 20)
        xs = [(torch.randn(3, 32000, generator=torch.Generator().manual_seed(100 + r)) * 0.1).to(dev) for r in range(world)]
        want = None
        for r in range(world):
            _, g = _local_grads(model, loss_fn, xs[r], 7 + r)
            want = g if want is None else {k: want[k] + g[k] for k in g}
        want = {k: v / world for k, v in want.items()}
        net = DataParallel(model)
        for rep in range(4):
            _, got = _local_grads(net, loss_fn, xs[rank], 7 + rank)
            torch.cuda.synchronize()
            got = {k.replace("module.", "", 1): v for k, v in got.items()}
            assert set(got) == set(want)
            for k in want:
                err = (got[k].float() - want[k].float()).abs().max().item()
                scale = want[k].abs().max().item() + 1e-8
                assert err <= 3e-3 * scale + 1e-7, f"rank {rank} step {rep} grad {k}: {err:.3g} vs scale {scale:.3g}"
        a = net._arena
        n_in = sum(1 for p in model.parameters() if p.grad is not None and a.buf.data_ptr() <= p.grad.data_ptr() < a.buf.data_ptr() + 4 * a.buf.numel())
        assert n_in == sum(1 for p in model.parameters() if p.grad is not None), "every gradient should live in the arena"
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def _switch_worker(rank, world, port, out_dir):
    """csrc/allreduce_mc.cu (multimem.ld_reduce / multimem.st through the NVSwitch) against NCCL's averaging all-reduce on
    the same data, over ragged ranges of one symmetric buffer"""
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from audio8_b200.parallel import SwitchAllReduce
        n = 3_000_000 + 64
        try:
            sw = SwitchAllReduce(n, dev, None)
        except Exception as e:  # no multicast on this box: nothing to test (the wrapper then uses NCCL)
            with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
                f.write(f"skip {e!r}")
            return
        g = torch.Generator(device=dev).manual_seed(10 + rank)
        src = torch.randn(n, device=dev, generator=g)
        want = src.clone()
        ranges = [(0, 4), (4, 1028), (1028, 2_000_000), (2_000_000, n)]
        for lo, hi in ranges:
            dist.all_reduce(want[lo:hi], op=dist.ReduceOp.AVG)
        sw.buf.copy_(src)
        torch.cuda.synchronize()
        dist.barrier()
        for lo, hi in ranges:
            sw.start(lo, hi).wait()
        torch.cuda.synchronize()
        err = (sw.buf - want).abs().max().item()
        assert err <= 1e-6 * want.abs().max().item() + 1e-7, f"rank {rank}: switch all-reduce differs from NCCL by {err:.3g}"
        every = [torch.empty(n, device=dev) for _ in range(world)]
        dist.all_gather(every, sw.buf)
        assert all(torch.equal(e, every[0]) for e in every), "ranks hold different results"
        # an empty range and a range that leaves the rest untouched
        before = sw.buf.clone()
        sw.start(128, 128).wait()
        torch.cuda.synchronize()
        assert torch.equal(before, sw.buf)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.gpu
def test_switch_allreduce_matches_nccl_world2(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2
    mp.spawn(_switch_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    notes = [open(tmp_path / f"ok{r}").read() for r in range(world)]
    if any(n.startswith("skip") for n in notes):
        pytest.skip(f"no NVSwitch multicast here: {notes[0]}")
    assert all(n == "ok" for n in notes)


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "switch"])
def test_arena_data_parallel_nccl_world2(tmp_path, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path), exchange), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


@pytest.mark.gpu
def test_arena_data_parallel_layerdrop_nccl_world2(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2
    mp.spawn(_layerdrop_worker, args=(world, _free_port(), str(tmp_path), "nccl"), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))
