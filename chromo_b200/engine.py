"""Engine: R replicas of one problem shape resident on one B200.

Thin object wrapper over the C ABI (include/chromo_b200.h); all arithmetic runs
in the CUDA kernels of chromo_b200/csrc.  Host arrays follow the reference's
layouts (fp64 [R,N,3], int64 [R,N,nb]).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import MOVE_DTYPE, NUM_MOVES, RNG_PHILOX, RNG_REPLAY, Shape, StepReport, check


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected array of shape {tuple(shape)}, got {a.shape}")
    return a


def _i64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.int64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected array of shape {tuple(shape)}, got {a.shape}")
    return a


def binding_free_energy_table(binders: Sequence[dict]) -> np.ndarray:
    """bind_F[b][Nm][s] = -ln sum_i C(Nm,i) C(Nn-Nm,s-i) exp(-(i e_mod + (s-i) e_nomod)),
    evaluated exactly as bead_binding_dE does (polymers.pyx:1493-1517: numpy
    arange/exp/sum/log and scipy.special.comb), once per (binder, marks, state)."""
    from scipy.special import comb
    S = max([b["sites_per_bead"] for b in binders] + [0])
    F = np.zeros((len(binders), S + 1, S + 1))
    for bi, b in enumerate(binders):
        Nn = b["sites_per_bead"]
        for Nm in range(S + 1):
            for s in range(S + 1):
                i = np.arange(s + 1)
                with np.errstate(divide="ignore"):
                    F[bi, Nm, s] = -np.log(np.sum(
                        comb(Nm, i) * comb(Nn - Nm, s - i) *
                        (np.exp(-(i * b["bind_energy_mod"] + (s - i) * b["bind_energy_no_mod"])))))
    return F


def host_register(a: np.ndarray) -> bool:
    """Page-lock a C-contiguous host array (`chromo_host_register`) so that host-array calls move it at link
    speed.  True: locked by this call (pair it with `host_unregister` before the array is freed); False: the
    memory is already page-locked (e.g. a view of torch pinned memory)."""
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("host_register needs a C-contiguous array")
    L = _lib.lib()
    rc = L.chromo_host_register(a.ctypes.data, a.nbytes)
    if rc == _lib.ERR_STATE:
        return False
    check(rc)
    return True


def host_unregister(a: np.ndarray) -> None:
    check(_lib.lib().chromo_host_unregister(a.ctypes.data))


class Engine:
    def __init__(self, n_replicas: int, num_beads: int, num_binders: int, *, grid: Optional[dict],
                 bead_vol: float, max_binders: int = -1, device: int = 0):
        """grid: dict(x_width, nx, y_width, ny, z_width, nz, confine_type, confine_length,
        vf_limit) as UniformDensityField.__init__ (fields.pyx:453-500), or None for NullField."""
        self.R, self.N, self.nb = int(n_replicas), int(num_beads), int(num_binders)
        s = Shape()
        s.n_replicas, s.num_beads, s.num_binders = self.R, self.N, self.nb
        if grid is not None:
            s.nx, s.ny, s.nz = int(grid["nx"]), int(grid["ny"]), int(grid["nz"])
            s.width[0], s.width[1], s.width[2] = grid["x_width"], grid["y_width"], grid["z_width"]
            ct = grid.get("confine_type", "")
            if ct not in _lib.CONFINE:
                raise ValueError("Confinement type " + str(ct) + " not found.")  # fields.pyx:198-201
            s.confine_type = _lib.CONFINE[ct]
            s.confine_length = grid.get("confine_length", 0.0)
            s.vf_limit = grid.get("vf_limit", 0.5)
            self.n_bins = s.nx * s.ny * s.nz
        else:
            self.n_bins = 0
        s.bead_vol = bead_vol
        s.max_binders = max_binders
        self._h = C.c_void_p()
        self._L = _lib.lib()
        check(self._L.chromo_ctx_create(C.byref(self._h), device, C.byref(s)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.chromo_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------- parameters
    def set_binders(self, binders: Sequence[dict], field_pref, e_intra, xpref):
        sites = _i64([b["sites_per_bead"] for b in binders], (self.nb,))
        F = _f64(binding_free_energy_table(binders))
        S = F.shape[1] - 1
        pref, e_intra, xpref = _f64(field_pref, (self.nb,)), _f64(e_intra, (self.nb,)), _f64(xpref, (self.nb, self.nb))
        check(self._L.chromo_set_binders(self._h, _lib.lptr(sites), _lib.dptr(pref), _lib.dptr(e_intra),
                                         _lib.dptr(xpref), _lib.dptr(F), S))
        self.bind_F = F

    def set_replica_params(self, chi=None, mu=None):
        chi = None if chi is None else _f64(np.broadcast_to(chi, (self.R,)))
        mu = None if mu is None else _f64(np.broadcast_to(mu, (self.R, self.nb)))
        check(self._L.chromo_set_replica_params(self._h, _lib.dptr(chi), _lib.dptr(mu)))

    def set_bond_params(self, eps_bend, eps_par, eps_perp, gamma, eta):
        arrs = [_f64(a) for a in (eps_bend, eps_par, eps_perp, gamma, eta)]
        n_sets = 1 if arrs[0].ndim == 1 else arrs[0].shape[0]
        for a in arrs:
            if a.size != n_sets * (self.N - 1):
                raise ValueError("bond parameter arrays must have N-1 entries per set")
        check(self._L.chromo_set_bond_params(self._h, n_sets, *[_lib.dptr(a) for a in arrs]))

    def set_twist_params(self, eps_twist=None, natural_twist=None):
        """SSTWLC: twist modulus and natural twist per bond ([N-1] or [R, N-1]); None switches the
        twist term off (polymers.pyx:2000, 2088-2090)."""
        if eps_twist is None:
            check(self._L.chromo_set_twist_params(self._h, 0, None, None))
            return
        a, b = _f64(eps_twist), _f64(natural_twist)
        n_sets = 1 if a.ndim == 1 else a.shape[0]
        if a.size != n_sets * (self.N - 1) or b.size != a.size:
            raise ValueError("twist parameter arrays must have N-1 entries per set")
        check(self._L.chromo_set_twist_params(self._h, n_sets, _lib.dptr(a), _lib.dptr(b)))

    def set_access_volumes(self, access_vol=None):
        a = None if access_vol is None else _f64(access_vol, (self.n_bins,))
        check(self._L.chromo_set_access_volumes(self._h, _lib.dptr(a)))

    # --------------------------------------------------------------- state
    def upload(self, r=None, t3=None, t2=None, states=None, mods=None, first=0, n=None):
        n = self.R - first if n is None else n
        sh3, shb = (n, self.N, 3), (n, self.N, self.nb)
        r = None if r is None else _f64(r).reshape(sh3)
        t3 = None if t3 is None else _f64(t3).reshape(sh3)
        t2 = None if t2 is None else _f64(t2).reshape(sh3)
        states = None if states is None else _i64(states).reshape(shb)
        mods = None if mods is None else _i64(mods).reshape(shb)
        check(self._L.chromo_upload_state(self._h, first, n, _lib.dptr(r), _lib.dptr(t3), _lib.dptr(t2),
                                          _lib.lptr(states), _lib.lptr(mods)))

    def download(self, first=0, n=None, want_states=True):
        n = self.R - first if n is None else n
        r = np.empty((n, self.N, 3))
        t3 = np.empty((n, self.N, 3))
        t2 = np.empty((n, self.N, 3))
        st = np.empty((n, self.N, self.nb), dtype=np.int64) if want_states else None
        check(self._L.chromo_download_state(self._h, first, n, _lib.dptr(r), _lib.dptr(t3), _lib.dptr(t2),
                                            _lib.lptr(st)))
        return r, t3, t2, st

    def download_into(self, r, t3, t2, states, first=0, n=None):
        n = self.R - first if n is None else n
        check(self._L.chromo_download_state(self._h, first, n, _lib.dptr(r), _lib.dptr(t3), _lib.dptr(t2),
                                            _lib.lptr(states)))

    def density(self, first=0, n=None):
        n = self.R - first if n is None else n
        d = np.empty((n, self.n_bins, self.nb + 1))
        check(self._L.chromo_download_density(self._h, first, n, _lib.dptr(d)))
        return d

    def upload_density(self, density, first=0):
        d = _f64(density)
        n = d.size // (self.n_bins * (self.nb + 1))
        check(self._L.chromo_upload_density(self._h, first, n, _lib.dptr(d)))

    # ------------------------------------------------------ full recompute
    def field_recompute(self, clamp=False):
        check(self._L.chromo_field_recompute(self._h, int(bool(clamp))))

    def field_energy(self):
        E = np.empty(self.R)
        sq = np.empty((self.R, self.nb))
        dbl = np.empty((self.R, self.nb), dtype=np.int64)
        ns = np.empty(self.R)
        check(self._L.chromo_field_energy(self._h, _lib.dptr(E), _lib.dptr(sq), _lib.lptr(dbl), _lib.dptr(ns)))
        return E, sq, dbl, ns

    def elastic_energy(self):
        E = np.empty(self.R)
        check(self._L.chromo_elastic_energy(self._h, _lib.dptr(E)))
        return E

    def chi_observable(self):
        P = np.empty(self.R)
        check(self._L.chromo_chi_observable(self._h, _lib.dptr(P)))
        return P

    # ----------------------------------------------------------------- RNG
    def srand(self, seeds):
        s = np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, dtype=np.uint32), (self.R,)))
        check(self._L.chromo_srand(self._h, _lib.uptr(s)))

    def numpy_seed(self, seeds):
        s = np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, dtype=np.uint32), (self.R,)))
        check(self._L.chromo_numpy_seed(self._h, _lib.uptr(s)))

    # ------------------------------------------------------------ hot path
    def set_moves(self, moves: np.ndarray):
        m = np.ascontiguousarray(moves, dtype=MOVE_DTYPE).reshape(self.R, NUM_MOVES)
        check(self._L.chromo_set_moves(self._h, m.ctypes.data_as(C.c_void_p)))

    def get_moves(self) -> np.ndarray:
        m = np.zeros((self.R, NUM_MOVES), dtype=MOVE_DTYPE)
        check(self._L.chromo_get_moves(self._h, m.ctypes.data_as(C.c_void_p)))
        return m

    def mc_sim(self, num_mc_steps: int, moves: Optional[np.ndarray] = None, mu_adjust_factor: float = 1.0,
               seed: int = 0, rng_mode: int = RNG_PHILOX, numpy_seeds=None):
        """mc_sim (mc_sim.pyx:26-103) for all replicas.  With `moves` given the
        call is synchronous and `moves` is updated in place."""
        mp = None
        if moves is not None:
            if moves.dtype != MOVE_DTYPE or not moves.flags.c_contiguous or moves.size != self.R * NUM_MOVES:
                raise ValueError("moves must be a C-contiguous [R,5] array of MOVE_DTYPE")
            mp = moves.ctypes.data_as(C.c_void_p)
        ns = None
        if numpy_seeds is not None:
            ns = np.ascontiguousarray(np.broadcast_to(np.asarray(numpy_seeds, dtype=np.uint32), (self.R,)))
        check(self._L.chromo_mc_sim(self._h, int(num_mc_steps), mp, float(mu_adjust_factor),
                                    int(seed) & 0xFFFFFFFFFFFFFFFF, int(rng_mode), _lib.uptr(ns)))

    def mc_sim_host(self, num_mc_steps: int, r, t3, t2, states, mods, moves: Optional[np.ndarray] = None,
                    mu_adjust_factor: float = 1.0, seed: int = 0, rng_mode: int = RNG_PHILOX, numpy_seeds=None,
                    n_chunks: int = 0):
        """mc_sim on HOST arrays in the reference's layouts, in place (r, t3, t2 [R,N,3] f64; states,
        mods [R,N,nb] int64): upload, kernel and download pipelined over replica chunks
        (chromo_mc_sim_host).  The arrays must be C-contiguous and of exactly these dtypes -- they are
        written through their own memory, no copies are made.  `mods=None`: the marks already on the
        device are current (mc_sim never modifies them)."""
        for name, a, dt in (("r", r, np.float64), ("t3", t3, np.float64), ("t2", t2, np.float64),
                            ("states", states, np.int64)) + ((("mods", mods, np.int64),) if mods is not None else ()):
            want = self.R * self.N * (3 if dt is np.float64 else self.nb)
            if not isinstance(a, np.ndarray) or a.dtype != dt or not a.flags.c_contiguous or a.size != want:
                raise ValueError(f"`{name}` must be a C-contiguous {np.dtype(dt).name} array of {want} elements")
        for name, a in (("r", r), ("t3", t3), ("t2", t2), ("states", states)):
            if not a.flags.writeable:
                raise ValueError(f"`{name}` is read-only: mc_sim writes its result into it")
        mp = None
        if moves is not None:
            if moves.dtype != MOVE_DTYPE or not moves.flags.c_contiguous or moves.size != self.R * NUM_MOVES:
                raise ValueError("moves must be a C-contiguous [R,5] array of MOVE_DTYPE")
            mp = moves.ctypes.data_as(C.c_void_p)
        ns = None
        if numpy_seeds is not None:
            ns = np.ascontiguousarray(np.broadcast_to(np.asarray(numpy_seeds, dtype=np.uint32), (self.R,)))
        check(self._L.chromo_mc_sim_host(self._h, int(num_mc_steps), mp, float(mu_adjust_factor),
                                         int(seed) & 0xFFFFFFFFFFFFFFFF, int(rng_mode), _lib.uptr(ns),
                                         _lib.dptr(r), _lib.dptr(t3), _lib.dptr(t2), _lib.lptr(states),
                                         _lib.lptr(mods) if mods is not None else None, int(n_chunks)))

    def set_table_capacity(self, cap: int = 0) -> int:
        """Slots of the per-replica shared-memory delta-density hash (0 = auto)."""
        out = C.c_int64(0)
        check(self._L.chromo_ctx_set_table_capacity(self._h, int(cap), C.byref(out)))
        return int(out.value)

    def set_warps_per_replica(self, warps: int = 0) -> int:
        """Warps that work on one replica in the production MC kernel (1 or 2; 0 = query)."""
        out = C.c_int64(0)
        check(self._L.chromo_ctx_set_warps_per_replica(self._h, int(warps), C.byref(out)))
        return int(out.value)

    def set_replicas_per_block(self, rpb: int = 0) -> int:
        """Replicas sharing one thread block of the MC kernel (1..7; -1 = automatic; 0 = query)."""
        out = C.c_int64(0)
        check(self._L.chromo_ctx_set_replicas_per_block(self._h, int(rpb), C.byref(out)))
        return int(out.value)

    def set_replica_offset(self, offset: int) -> None:
        """Global index of this context's replica 0: the production streams are keyed by
        (seed, global replica, attempt), so the shards of one ensemble draw different numbers."""
        check(self._L.chromo_ctx_set_replica_offset(self._h, int(offset)))

    def set_move_order(self, order) -> None:
        """Order of the move types within one MC step (a permutation of the move ids; the reference walks its
        controller list, mc_sim.pyx:92-103)."""
        a = np.ascontiguousarray(order, dtype=np.int32).reshape(NUM_MOVES)
        check(self._L.chromo_ctx_set_move_order(self._h, a.ctypes.data_as(C.POINTER(C.c_int32))))

    def set_batch_size(self, batch: int) -> None:
        """Attempts prepared at once by the production kernels (1..32; test knob, results do not depend on it)."""
        check(self._L.chromo_ctx_set_batch_size(self._h, int(batch)))

    def set_detailed_nucleosomes(self, consts20=None) -> None:
        """DetailedChromatin: the 20 nucleosome constants of `util.nucleo_geom.nucleosome_constants(bp_wrap)`
        (None switches the entry / exit geometry off); set the twist parameters first."""
        if consts20 is None:
            check(self._L.chromo_set_detailed_nucleosomes(self._h, None))
            return
        a = _f64(np.asarray(consts20, dtype=float).reshape(20))
        check(self._L.chromo_set_detailed_nucleosomes(self._h, _lib.dptr(a)))

    def set_fast_field(self, n_points: int) -> None:
        """fast_field = 1 of the reference's UniformDensityField: sub-bin quantised binning in the dE path
        (`n_points` sub-bins per voxel edge; 0 switches it off)."""
        check(self._L.chromo_ctx_set_fast_field(self._h, int(n_points)))

    def rng_counters(self) -> np.ndarray:
        """Attempts made so far per replica = position of its production random stream."""
        out = np.zeros(self.R, dtype=np.uint64)
        check(self._L.chromo_get_rng_counters(self._h, 0, self.R, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out

    def set_rng_counters(self, counters) -> None:
        a = np.ascontiguousarray(counters, dtype=np.uint64)
        if a.shape != (self.R,):
            raise ValueError(f"counters must have shape ({self.R},)")
        check(self._L.chromo_set_rng_counters(self._h, 0, self.R, a.ctypes.data_as(C.POINTER(C.c_uint64))))

    def sync(self):
        check(self._L.chromo_ctx_sync(self._h))

    def last_attempts(self) -> int:
        return int(self._L.chromo_last_attempts(self._h))

    def last_algo_bytes(self) -> int:
        return int(self._L.chromo_last_algo_bytes(self._h))

    def stream(self) -> int:
        return int(self._L.chromo_ctx_stream(self._h) or 0)

    def bytes(self) -> int:
        return int(self._L.chromo_ctx_bytes(self._h))

    def mc_step(self, replica: int, move: int, amp_move: float, amp_bead: int, mu_adjust_factor: float = 1.0,
                rng_mode: int = RNG_REPLAY, seed: int = 0, force_accept: int = -1):
        """One instrumented mc_step (mc_sim.pyx:106-182) of one replica."""
        rep = StepReport()
        inds = np.zeros(self.N, dtype=np.int64)
        rows = np.zeros((self.N, 9 + self.nb))
        tcap = max(1, min(self.n_bins, 16 * self.N))
        touched = np.zeros(tcap, dtype=np.int64)
        dtrial = np.zeros((tcap, self.nb + 1))
        check(self._L.chromo_mc_step(self._h, replica, move, float(amp_move), int(amp_bead),
                                     float(mu_adjust_factor), rng_mode, int(seed), int(force_accept),
                                     C.byref(rep), _lib.lptr(inds), self.N, _lib.dptr(rows), self.N,
                                     _lib.lptr(touched), _lib.dptr(dtrial), tcap))
        n, nt = rep.n_inds, rep.n_touched
        return dict(inds=inds[:n].copy(), rows=rows[:n].copy(), touched=touched[:nt].copy(),
                    dtrial=dtrial[:nt].copy(), dE_poly=rep.dE_poly, dE_field=rep.dE_field, u=rep.u,
                    accepted=bool(rep.accepted), passes=rep.passes)
