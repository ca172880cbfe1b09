"""The C-ABI library loads on a CPU-only box and exports every symbol include/audio8_b200.h declares, with the
argument count the ctypes table binds (no compute calls here)."""
import os
import re

from audio8_b200 import _lib

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "audio8_b200.h")


def _declared():
    h = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(?:int|void|size_t|int64_t|const char\*)\s+(a8_\w+)\s*\(([^;]*?)\)\s*;", h, flags=re.S):
        out[name] = 0 if args.strip() == "void" else len(args.split(","))
    return out


def test_header_symbols_exported_and_bound():
    lib = _lib.load()
    decl = _declared()
    assert len(decl) >= 29
    for name, nargs in decl.items():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, ctypes table {len(_lib.SIGNATURES[name][1])}"
    assert set(_lib.SIGNATURES) == set(decl)
    assert lib.a8_version() == 3


def test_struct_layout_matches_header():
    import ctypes as C
    assert C.sizeof(_lib.Operand) == 168
    assert _lib.Gemm.c.offset == 2 * 168 + 32
    assert C.sizeof(_lib.Gemm) % 8 == 0


def test_no_cpu_fallback():
    """the product backend refuses CPU tensors instead of silently computing something else"""
    import pytest
    import torch
    from audio8_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    be = ops.CudaBackend()
    with pytest.raises(AssertionError):
        be.layernorm_fwd(torch.zeros(4, 8, dtype=torch.bfloat16), torch.ones(8), torch.zeros(8), 1e-5)
