"""The C-ABI shared library: it loads, and exports every symbol that
include/chromo_b200.h declares (no compute calls -- this runs without a GPU)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "chromo_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(chromo_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    from chromo_b200 import _lib
    assert _declared() == sorted(_lib.SYMBOLS)


def test_cuda_library_exports_every_symbol():
    lib = ROOT / "chromo_b200" / "libchromo_b200.so"
    import shutil
    if shutil.which("nvcc") or Path("/usr/local/cuda/bin/nvcc").exists():
        import __graft_entry__ as g
        g.build_cuda()  # no-op when the in-tree build is up to date
    L = ctypes.CDLL(str(lib))
    for s in _declared():
        assert hasattr(L, s), s
    L.chromo_version.restype = ctypes.c_int
    assert L.chromo_version() >= 100


def test_move_state_layout():
    from chromo_b200._lib import MOVE_DTYPE, Shape, StepReport
    # offsets of chromo_move_state in include/chromo_b200.h
    want = dict(amp_move=0, move_amp_lo=8, move_amp_hi=16, bead_amp_lo=24, bead_amp_hi=32,
                acceptance_rate=40, alpha=48, num_attempt=56, num_success=64, amp_bead=72,
                num_per_cycle=76, move_on=80, controller=84)
    for k, off in want.items():
        assert MOVE_DTYPE.fields[k][1] == off, k
    assert MOVE_DTYPE.itemsize == 88
    assert ctypes.sizeof(Shape) == 112 and ctypes.sizeof(StepReport) == 48


def test_fails_loudly_without_library(tmp_path, monkeypatch):
    from chromo_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "missing.so")
    monkeypatch.setattr(_lib, "_LIB", None)
    with pytest.raises(_lib.ChromoError, match="no CPU fallback"):
        _lib.lib()


def test_no_gpu_is_an_error_not_a_fallback():
    """Without a CUDA device the real library refuses to create a context."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from chromo_b200 import _lib
    _lib._LIB = None
    try:
        from chromo_b200.engine import Engine
        with pytest.raises(_lib.ChromoError):
            Engine(1, 10, 1, grid=None, bead_vol=1.0)
        # the context-free entry points (coarse-grain / refine) refuse as well
        import numpy as np
        import chromo_b200.util.rediscretize as rd
        with pytest.raises(_lib.ChromoError, match="no CUDA device|no CPU fallback"):
            rd.coarse_grain_ensemble(np.zeros((1, 10, 3)), np.ones((1, 10, 3)), None, None, 2)
        with pytest.raises(_lib.ChromoError, match="no CUDA device|no CPU fallback"):
            rd.get_refined_path(np.zeros((3, 3)), 10, 1.0)
    finally:
        _lib._LIB = None
