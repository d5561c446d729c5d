import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "oracle", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

EMU_LIB = ROOT / "tests" / "host_emu" / "libchromo_emu.so"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the reference build in oracle/_ref")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()
    return oracle


def _emu_stale():
    if not EMU_LIB.exists():
        return True
    t = EMU_LIB.stat().st_mtime
    deps = list((ROOT / "chromo_b200" / "csrc").glob("*.cu*")) + [ROOT / "include" / "chromo_b200.h",
                                                                ROOT / "tests" / "host_emu" / "cuda_emu.h"]
    return any(d.stat().st_mtime > t for d in deps)


@pytest.fixture(scope="session")
def emu_lib():
    """The lockstep-warp CPU emulation of the kernels (tests/host_emu): a
    debugging twin of the CUDA library, used only by the not-gpu tests."""
    if _emu_stale():
        subprocess.run([str(ROOT / "tests" / "host_emu" / "build.sh")], check=True)
    return EMU_LIB


BACKENDS = ["emu", pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def backend(request, emu_lib):
    """Binds chromo_b200 to the emulated twin ("emu", CPU box) or to the real
    sm_100a library ("cuda", GPU box) for the duration of one test."""
    from chromo_b200 import _lib
    if request.param == "emu":
        import devlib
        devlib.use_library(emu_lib)
    else:
        _lib._LIB = None
        _lib.lib()
    yield request.param
    _lib._LIB = None


@pytest.fixture
def cuda_backend():
    from chromo_b200 import _lib
    _lib._LIB = None
    _lib.lib()
    yield "cuda"
    _lib._LIB = None
