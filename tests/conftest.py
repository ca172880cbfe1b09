import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def emu_backend():
    """Inject the pure-PyTorch ABI emulation (tests/emu.py) for CPU tests of the host-side logic."""
    from audio8_b200 import ops
    import emu
    prev = ops._BACKEND
    ops.set_backend(emu.EmuOps())
    yield ops._BACKEND
    ops.set_backend(prev)


def pytest_sessionfinish(session, exitstatus):
    """every gradient comparison of a GPU session, with its measured cosine / rel-L2 and the tolerance it was held to,
    as a markdown table under gpurun_out/ (copied to profiles/r02_parity.md)"""
    try:
        import torch
        import model_cases as mc
        if not mc.PARITY_LOG or not torch.cuda.is_available():
            return
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.md"), "w") as f:
            f.write("| case | tensor | cosine | rel-L2 | held to (cos >= / rel <=) | bf16-storage oracle vs fp32 oracle "
                    "(cosine / rel-L2) |\n|---|---|---:|---:|---|---|\n")
            for case, what, cos, rel, cmin, rmax, fl in mc.PARITY_LOG:
                fls = f"{fl[0]:.6f} / {fl[1]:.5f}" if fl else ""
                f.write(f"| {case} | {what} | {cos:.6f} | {rel:.5f} | {cmin} / {rmax} | {fls} |\n")
            if mc.VQ_LOG:
                f.write("\n| case | VQ code-index flips vs the oracle's arg-max | entries |\n|---|---:|---:|\n")
                for case, flips, n in mc.VQ_LOG:
                    f.write(f"| {case} | {flips} | {n} |\n")
    except Exception as e:  # reporting only
        print("parity report not written:", e)
