"""Optimizer side of the step: `audio8_b200.optim.FusedAdamW` / `clip_grad_norm_` against
`torch.optim.AdamW` + `torch.nn.utils.clip_grad_norm_` (what pretrain.py:182-184 / train.py:323-325 run) on identical
parameters and gradients: parameters and optimizer state to 1e-6 relative after several steps; the returned norm to
1e-5 (ours is accumulated in double precision; torch's fp32 sum over ~6e5 elements is itself ~4e-6 off the exact value)."""
import pytest
import torch


def _params(device, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(768, 768), (3072,), (513, 37), (1, 1, 128), (7,), (70001,), (256, 3, 5)]  # odd sizes: scalar tails, 3 chunks
    return [torch.nn.Parameter((torch.randn(s, generator=g) * 0.1).to(device)) for s in shapes]


def _run(device, clip, grad_scale, steps=4, skip_grad=False):
    from audio8_b200.optim import FusedAdamW, clip_grad_norm_
    ours, ref = _params(device), _params(device)
    kw = dict(lr=3e-3, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.01)
    o1 = FusedAdamW(ours, **kw)
    o2 = torch.optim.AdamW(ref, **kw)
    bf = torch.empty(ours[0].numel(), dtype=torch.bfloat16, device=device)
    o1.register_operand_copy(ours[0], bf)
    g = torch.Generator().manual_seed(5)
    for it in range(steps):
        for a, b in zip(ours, ref):
            gr = (torch.randn(a.shape, generator=g) * (3.0 if it == 1 else 0.01)).to(device)
            a.grad, b.grad = gr.clone(), gr.clone()
        if skip_grad:
            ours[4].grad = ref[4].grad = None  # a parameter without a gradient is left untouched
        if grad_scale != 1.0:
            for b in ref:
                if b.grad is not None:
                    b.grad.mul_(grad_scale)
        if clip is not None:
            n_ref = torch.nn.utils.clip_grad_norm_(ref, clip)
        if it % 2 == 0 or clip is None:
            o1.step(clip=clip, grad_scale=grad_scale)  # fused unscale + clip + update (gradients never rewritten)
            n_ours = o1.last_grad_norm
        else:  # the trainers' unmodified call pattern (train.py:323-325): scale in place, clip in place, plain step
            if grad_scale != 1.0:
                o1.scale_grads(grad_scale)  # eight_mile OptimizerManager.scale_grads
            n_ours = clip_grad_norm_(ours, clip)
            o1.step()
        o2.step()
        if clip is not None:
            assert abs(float(n_ours) - float(n_ref)) <= 1e-5 * float(n_ref), (it, float(n_ours), float(n_ref))
    for i, (a, b) in enumerate(zip(ours, ref)):
        err = (a.detach() - b.detach()).abs().max().item()
        assert err <= 1e-6 * b.detach().abs().max().item() + 1e-9, f"param {i}: {err:.3g}"
        if skip_grad and i == 4:
            continue
        for key in ("exp_avg", "exp_avg_sq"):
            sa, sb = o1.state[a][key], o2.state[b][key]
            tol = 2e-6 if clip is None else 1e-5  # with clipping the coefficient inherits torch's fp32 norm error
            assert (sa - sb).abs().max().item() <= tol * sb.abs().max().item() + 1e-12, (i, key)
    assert torch.equal(bf.view(ours[0].shape), ours[0].detach().to(torch.bfloat16)), "bf16 operand copy not refreshed"
    # state_dict round trip with torch's optimizer
    o3 = torch.optim.AdamW(_params(device), **kw)
    o3.load_state_dict(o1.state_dict())


CASES = [(None, 1.0, False), (0.5, 1.0, False), (25.0, 1.0, True), (1.0, 0.125, False)]


@pytest.mark.parametrize("clip,grad_scale,skip", CASES)
def test_fused_adamw_host_logic_cpu(clip, grad_scale, skip, emu_backend):
    _run("cpu", clip, grad_scale, skip_grad=skip)


@pytest.mark.gpu
@pytest.mark.parametrize("clip,grad_scale,skip", CASES)
def test_fused_adamw_cuda(clip, grad_scale, skip):
    _run("cuda", clip, grad_scale, skip_grad=skip)


@pytest.mark.gpu
def test_fused_adamw_on_model_cuda():
    """the whole base model's parameter list (208 tensors, 95 M elements) after a real backward: same update as torch"""
    import numpy as np
    from audio8_b200 import wav2vec2 as W
    from audio8_b200.optim import FusedAdamW
    torch.manual_seed(0)
    np.random.seed(0)
    model = W.create_model(num_layers=2).cuda().train()
    loss_fn = W.create_loss(640, 100)
    x = torch.randn(2, 32000, device="cuda") * 0.1
    loss_fn(model, x).backward()
    ref = [torch.nn.Parameter(p.detach().clone()) for p in model.parameters()]
    for r, p in zip(ref, model.parameters()):
        r.grad = p.grad.detach().clone() if p.grad is not None else None
    o1 = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.01)
    o2 = torch.optim.AdamW(ref, lr=1e-3, weight_decay=0.01)
    n_ref = torch.nn.utils.clip_grad_norm_(ref, 1.0)
    o1.step(clip=1.0)
    o2.step()
    assert abs(float(o1.last_grad_norm) - float(n_ref)) <= 1e-5 * float(n_ref)
    for p, r in zip(model.parameters(), ref):
        assert (p.detach() - r.detach()).abs().max().item() <= 1e-6 * r.detach().abs().max().item() + 1e-9
