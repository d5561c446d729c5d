"""Test / development hook: bind chromo_b200 to a specific build of the C ABI -- the CPU emulation of the
kernels (tests/host_emu) on a GPU-less box, or an A/B build of the CUDA library (tools/build_variant.sh).
Lives outside the product package on purpose: chromo_b200 itself only ever loads its in-tree
libchromo_b200.so and fails loudly when it is missing."""
import ctypes


def use_library(path):
    from chromo_b200 import _lib
    _lib._LIB = _lib._declare(ctypes.CDLL(str(path)))
    return _lib._LIB
