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
