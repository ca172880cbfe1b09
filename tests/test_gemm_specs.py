"""GEMM descriptors (conv windows, tap shifts, groups, heads) checked on CPU through the ABI emulation, and on
the GPU through the tcgen05 kernel, against plain PyTorch fp32 math."""
import pytest
import torch

import gemm_cases


@pytest.mark.parametrize("name", sorted(gemm_cases.ALL_CASES))
def test_spec_semantics_cpu(name, emu_backend):
    gemm_cases.run_case(name, "cpu", emu_backend)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(gemm_cases.ALL_CASES))
def test_tcgen05_gemm(name):
    from audio8_b200 import ops
    gemm_cases.run_case(name, "cuda", ops.backend())
