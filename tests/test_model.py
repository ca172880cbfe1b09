"""End-to-end parity of the drop-in modules (see tests/model_cases.py for cases and tolerances)."""
import pytest

import model_cases


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_pretrain_host_logic_cpu(mode, emu_backend):
    model_cases.run_pretrain_case("cpu", mode)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_acoustic_host_logic_cpu(mode, emu_backend):
    model_cases.run_acoustic_case("cpu", mode)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_pretrain_cuda(mode):
    model_cases.run_pretrain_case("cuda", mode)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_acoustic_cuda(mode):
    model_cases.run_acoustic_case("cuda", mode)


@pytest.mark.gpu
def test_cuda_graph_replay_matches_eager():
    model_cases.run_graph_case()


def test_pretrain_generic_host_logic_cpu(emu_backend):
    """the fixture-free parity driver itself, on CPU through the ABI emulation (helper-thread draws included)"""
    cfg = dict(d_model=128, num_heads=2, num_layers=2, d_ff=256, final_dim=64, num_vq_vars=24, num_vq_groups=2)
    model_cases.run_pretrain_generic("cpu", cfg, B=2, L=8000, K=10)


def test_pretrain_split_encoder_host_logic_cpu(emu_backend):
    """the encoder as two autograd nodes (what deep stacks use so that DDP can overlap the all-reduce)"""
    cfg = dict(d_model=128, num_heads=2, num_layers=2, d_ff=256, final_dim=64, num_vq_vars=24, num_vq_groups=2)
    model_cases.run_pretrain_generic("cpu", cfg, B=2, L=8000, K=10, split_min=2,
                                     check_grads=("mask_emb", "encoder.transformer.encoders.0.ffn.0.layer.weight",
                                                  "encoder.transformer.encoders.1.self_attn.w_O.layer.weight",
                                                  "encoder.ln.weight", "encoder.pos_conv.conv.1.weight_v"))


@pytest.mark.gpu
def test_pretrain_large_config_cuda():
    """wav2vec2-large widths (d=1024, 16 heads, d_ff=4096; BASELINE configs[3]) at a short crop, 2 layers"""
    cfg = dict(d_model=1024, num_heads=16, num_layers=2, d_ff=4096, final_dim=256, num_vq_vars=320, num_vq_groups=2)
    model_cases.run_pretrain_generic("cuda", cfg, B=2, L=16000, K=20)


@pytest.mark.gpu
def test_pretrain_full_size_cuda():
    """BASELINE configs[1] at FULL size (base model, B=6 x 15 s, K=100) against the CPU oracle: loss and a sample of
    gradients (about half a minute of host time for the oracle's fp32 forward+backward)."""
    cfg = dict(d_model=768, num_heads=12, num_layers=12, final_dim=256, num_vq_vars=320, num_vq_groups=2)
    ours, ref, vq = model_cases.run_pretrain_generic("cuda", cfg, B=6, L=240000, K=100)
    assert vq >= 0.95


def test_acoustic_generic_host_logic_cpu(emu_backend):
    cfg = dict(d_model=128, num_heads=2, num_layers=2, d_ff=256)
    model_cases.run_acoustic_generic("cpu", cfg, V=32, B=3, L=12000, S=8,
                                     check_grads=("proj.weight", "encoder.mask_emb",
                                                  "encoder.encoder.transformer.encoders.0.ffn.3.layer.weight"))


@pytest.mark.gpu
def test_acoustic_full_size_cuda():
    """BASELINE configs[2] per GPU at FULL size: wav2vec2-base CTC fine-tune (char vocab 32), B=8 x 15 s ragged,
    150-char targets, frozen feature encoder, time + channel masks, against the CPU oracle"""
    cfg = dict(d_model=768, num_heads=12, num_layers=12)
    model_cases.run_acoustic_generic("cuda", cfg, V=32, B=8, L=240000, S=150)


def test_activations_released_at_backward_cpu(emu_backend):
    """a Function's saved activations are dropped when its backward starts (a loss tensor kept across steps must not pin
    them), so a second backward through the same graph fails loudly instead of silently reusing freed state"""
    import torch
    from audio8_b200 import functional as Fn
    x = torch.randn(5, 16, requires_grad=True)
    w = torch.randn(8, 16, requires_grad=True)
    y = Fn.linear(x, w, None, out_f32=True)
    ctx = y.grad_fn
    assert ctx.saved is not None
    y.sum().backward(retain_graph=True)
    assert ctx.saved is None
    with pytest.raises(RuntimeError, match="second time"):
        y.sum().backward()
