"""ctypes binding of the C ABI in include/chromo_b200.h.

The shared library is built IN-TREE by `__graft_entry__.build()` (nvcc, sm_100a)
as chromo_b200/libchromo_b200.so.  There is no CPU fallback: if the library is
missing, or no CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libchromo_b200.so"

NUM_MOVES = 5
MOVE_NAMES = ("crank_shaft", "end_pivot", "slide", "tangent_rotation", "change_binding_state")
MOVE_ID = {n: i for i, n in enumerate(MOVE_NAMES)}
CONFINE = {"": 0, "Spherical": 1, "Cubical": 2}
RNG_PHILOX, RNG_REPLAY = 0, 1

# numpy mirror of `chromo_move_state` (include/chromo_b200.h)
MOVE_DTYPE = np.dtype([
    ("amp_move", "<f8"), ("move_amp_lo", "<f8"), ("move_amp_hi", "<f8"),
    ("bead_amp_lo", "<f8"), ("bead_amp_hi", "<f8"), ("acceptance_rate", "<f8"),
    ("alpha", "<f8"), ("num_attempt", "<i8"), ("num_success", "<i8"),
    ("amp_bead", "<i4"), ("num_per_cycle", "<i4"), ("move_on", "<i4"), ("controller", "<i4"),
], align=True)
assert MOVE_DTYPE.itemsize == 88


class Shape(C.Structure):
    _fields_ = [
        ("n_replicas", C.c_int64), ("num_beads", C.c_int64), ("num_binders", C.c_int64),
        ("nx", C.c_int64), ("ny", C.c_int64), ("nz", C.c_int64),
        ("width", C.c_double * 3), ("confine_type", C.c_int32), ("confine_length", C.c_double),
        ("vf_limit", C.c_float), ("bead_vol", C.c_double), ("max_binders", C.c_int64),
    ]


class StepReport(C.Structure):
    _fields_ = [
        ("n_inds", C.c_int64), ("n_touched", C.c_int64), ("dE_poly", C.c_double),
        ("dE_field", C.c_double), ("u", C.c_double), ("accepted", C.c_int32), ("passes", C.c_int32),
    ]


class ChromoError(RuntimeError):
    pass


_LIB = None
_pd = C.POINTER(C.c_double)
_pl = C.POINTER(C.c_int64)
_pu = C.POINTER(C.c_uint32)
_vp = C.c_void_p

# every symbol declared in include/chromo_b200.h
SYMBOLS = [
    "chromo_last_error", "chromo_version", "chromo_ctx_create", "chromo_ctx_destroy",
    "chromo_ctx_sync", "chromo_ctx_stream", "chromo_ctx_bytes", "chromo_ctx_set_table_capacity",
    "chromo_ctx_set_warps_per_replica", "chromo_ctx_set_replicas_per_block", "chromo_ctx_set_replica_offset", "chromo_ctx_set_batch_size", "chromo_ctx_set_move_order", "chromo_host_register", "chromo_host_unregister", "chromo_ctx_set_fast_field", "chromo_set_detailed_nucleosomes",
    "chromo_get_rng_counters", "chromo_set_rng_counters", "chromo_set_binders",
    "chromo_set_replica_params", "chromo_set_bond_params", "chromo_set_twist_params", "chromo_set_access_volumes",
    "chromo_upload_state", "chromo_download_state", "chromo_download_density",
    "chromo_upload_density", "chromo_field_recompute", "chromo_field_energy",
    "chromo_elastic_energy", "chromo_chi_observable", "chromo_exchange_init", "chromo_exchange_observable",
    "chromo_exchange_step", "chromo_exchange_state", "chromo_srand", "chromo_numpy_seed",
    "chromo_mc_sim", "chromo_mc_sim_host", "chromo_get_moves", "chromo_set_moves", "chromo_last_attempts", "chromo_last_algo_bytes",
    "chromo_mc_step",
    "chromo_cg_num_beads", "chromo_cg_chromatin", "chromo_refined_num_points", "chromo_refined_num_draws",
    "chromo_refine_path", "chromo_enforce_spherical_confinement",
]


def _declare(L):
    L.chromo_last_error.restype = C.c_char_p
    L.chromo_ctx_create.argtypes = [C.POINTER(_vp), C.c_int, C.POINTER(Shape)]
    L.chromo_ctx_destroy.argtypes = [_vp]
    L.chromo_ctx_sync.argtypes = [_vp]
    L.chromo_ctx_stream.argtypes = [_vp]
    L.chromo_ctx_stream.restype = _vp
    L.chromo_ctx_bytes.argtypes = [_vp]
    L.chromo_ctx_bytes.restype = C.c_int64
    L.chromo_ctx_set_table_capacity.argtypes = [_vp, C.c_int64, _pl]
    L.chromo_ctx_set_warps_per_replica.argtypes = [_vp, C.c_int64, _pl]
    L.chromo_ctx_set_replicas_per_block.argtypes = [_vp, C.c_int64, _pl]
    L.chromo_ctx_set_replica_offset.argtypes = [_vp, C.c_int64]
    L.chromo_ctx_set_batch_size.argtypes = [_vp, C.c_int64]
    L.chromo_ctx_set_fast_field.argtypes = [_vp, C.c_int64]
    L.chromo_ctx_set_move_order.argtypes = [_vp, C.POINTER(C.c_int32)]
    L.chromo_host_register.argtypes = [_vp, C.c_uint64]
    L.chromo_host_unregister.argtypes = [_vp]
    L.chromo_set_detailed_nucleosomes.argtypes = [_vp, _pd]
    L.chromo_get_rng_counters.argtypes = [_vp, C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]
    L.chromo_set_rng_counters.argtypes = [_vp, C.c_int64, C.c_int64, C.POINTER(C.c_uint64)]
    L.chromo_set_binders.argtypes = [_vp, _pl, _pd, _pd, _pd, _pd, C.c_int64]
    L.chromo_set_replica_params.argtypes = [_vp, _pd, _pd]
    L.chromo_set_bond_params.argtypes = [_vp, C.c_int64, _pd, _pd, _pd, _pd, _pd]
    L.chromo_set_twist_params.argtypes = [_vp, C.c_int64, _pd, _pd]
    L.chromo_set_access_volumes.argtypes = [_vp, _pd]
    L.chromo_upload_state.argtypes = [_vp, C.c_int64, C.c_int64, _pd, _pd, _pd, _pl, _pl]
    L.chromo_download_state.argtypes = [_vp, C.c_int64, C.c_int64, _pd, _pd, _pd, _pl]
    L.chromo_download_density.argtypes = [_vp, C.c_int64, C.c_int64, _pd]
    L.chromo_upload_density.argtypes = [_vp, C.c_int64, C.c_int64, _pd]
    L.chromo_field_recompute.argtypes = [_vp, C.c_int]
    L.chromo_field_energy.argtypes = [_vp, _pd, _pd, _pl, _pd]
    L.chromo_elastic_energy.argtypes = [_vp, _pd]
    L.chromo_chi_observable.argtypes = [_vp, _pd]
    L.chromo_exchange_init.argtypes = [_vp, _pd, C.c_int64, C.c_int64, C.c_int64]
    L.chromo_exchange_observable.argtypes = [_vp, C.c_void_p]
    L.chromo_exchange_step.argtypes = [_vp, C.c_void_p, C.c_int64, C.c_uint64]
    L.chromo_exchange_state.argtypes = [_vp, C.POINTER(C.c_int32), _pd, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.chromo_srand.argtypes = [_vp, _pu]
    L.chromo_numpy_seed.argtypes = [_vp, _pu]
    L.chromo_mc_sim.argtypes = [_vp, C.c_int64, _vp, C.c_double, C.c_uint64, C.c_int, _pu]
    L.chromo_mc_sim_host.argtypes = [_vp, C.c_int64, _vp, C.c_double, C.c_uint64, C.c_int, _pu,
                                     _pd, _pd, _pd, _pl, _pl, C.c_int64]
    L.chromo_get_moves.argtypes = [_vp, _vp]
    L.chromo_set_moves.argtypes = [_vp, _vp]
    L.chromo_last_attempts.argtypes = [_vp]
    L.chromo_last_attempts.restype = C.c_int64
    L.chromo_last_algo_bytes.argtypes = [_vp]
    L.chromo_last_algo_bytes.restype = C.c_int64
    L.chromo_mc_step.argtypes = [_vp, C.c_int64, C.c_int, C.c_double, C.c_int64, C.c_double, C.c_int,
                                 C.c_uint64, C.c_int, C.POINTER(StepReport), _pl, C.c_int64, _pd,
                                 C.c_int64, _pl, _pd, C.c_int64]
    L.chromo_cg_num_beads.argtypes = [C.c_int64, C.c_int64]
    L.chromo_cg_num_beads.restype = C.c_int64
    L.chromo_cg_chromatin.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_double, _pd, _pd,
                                      _pl, _pl, _pd, _pd, _pd, _pl, _pl, _pd]
    L.chromo_refined_num_points.argtypes = [C.c_int64, C.c_int64]
    L.chromo_refined_num_points.restype = C.c_int64
    L.chromo_refined_num_draws.argtypes = [C.c_int64, C.c_int64]
    L.chromo_refined_num_draws.restype = C.c_int64
    L.chromo_refine_path.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double, _pd, _pd, C.c_uint64,
                                     C.c_double, C.c_int, _pd, _pd, _pd]
    L.chromo_enforce_spherical_confinement.argtypes = [C.c_int, C.c_int64, C.c_int64, _pd, C.c_double, _pd]
    return L


def lib():
    """The CUDA library.  Fails loudly when it has not been built."""
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise ChromoError(
                f"{LIB_PATH} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()'). "
                "chromo_b200 has no CPU fallback.")
        _LIB = _declare(C.CDLL(str(LIB_PATH)))
    return _LIB


ERR_ARG, ERR_CUDA, ERR_STATE = -1, -2, -3  # include/chromo_b200.h:41-43


def check(rc):
    if rc != 0:
        msg = lib().chromo_last_error()
        raise ChromoError(f"chromo_b200 error {rc}: {msg.decode() if msg else '?'}")


def dptr(a):
    return a.ctypes.data_as(_pd) if a is not None else None


def lptr(a):
    return a.ctypes.data_as(_pl) if a is not None else None


def uptr(a):
    return a.ctypes.data_as(_pu) if a is not None else None
