"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference modules in the dev container.

`/root/reference` exists only in the dev container; nothing that runs on the GPU box may call this.
`audio8/__init__.py` star-imports `data.py` -> `soundfile` (absent), so a bare package object whose
`__path__` points at the reference directory is pre-seeded instead.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def load_reference():
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "audio8")):
        raise RuntimeError("/root/reference not present (GPU box?) — golden fixtures must be used instead")
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (here, os.path.join(here, "shim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "audio8" not in sys.modules or not hasattr(sys.modules["audio8"], "__path__"):
        pkg = types.ModuleType("audio8")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "audio8")]
        sys.modules["audio8"] = pkg
    import audio8.wav2vec2 as w2v  # noqa: E402
    import audio8.ctc as ctc  # noqa: E402
    return w2v, ctc
