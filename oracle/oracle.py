"""ctypes front-end of the CPU oracle (oracle/chromo_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg -- never by the product package chromo_b200/.

A problem is described by a plain `spec` dict (see `make_spec`) so that the
same inputs can be fed to (a) this oracle, (b) the reference's own Cython build
in oracle/_ref (see `ref_objects`), and (c) the CUDA product.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_LIB = None

MOVE_NAMES = ["crank_shaft", "end_pivot", "slide", "tangent_rotation", "change_binding_state"]
CONFINE = {"": 0, "Spherical": 1, "Cubical": 2}

_pd = C.POINTER(C.c_double)
_pl = C.POINTER(C.c_int64)


class GlibcRand(C.Structure):
    _fields_ = [("r", C.c_int32 * 34), ("f", C.c_int), ("b", C.c_int), ("alt", C.c_void_p)]


class Philox(C.Structure):
    _fields_ = [("k0", C.c_uint32), ("k1", C.c_uint32), ("rep", C.c_uint32), ("pos", C.c_uint32),
                ("attempt", C.c_uint64), ("next_attempt", C.c_uint64), ("blk", C.c_uint32 * 4)]


class Detailed(C.Structure):
    _fields_ = [("t3_local", C.c_double * 3), ("t2_local", C.c_double * 3), ("r_enter_unit", C.c_double * 3),
                ("r_enter_norm", C.c_double), ("r_exit_unit", C.c_double * 3), ("r_exit_norm", C.c_double),
                ("a3", C.c_double * 3), ("a1", C.c_double * 3)]


class MT19937(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("pos", C.c_int)]


class Move(C.Structure):
    _fields_ = [
        ("move_on", C.c_int64), ("num_per_cycle", C.c_int64), ("amp_move", C.c_double),
        ("amp_bead", C.c_int64), ("num_attempt", C.c_int64), ("num_success", C.c_int64),
        ("acceptance_rate", C.c_double), ("alpha", C.c_double),
        ("move_amp_lo", C.c_double), ("move_amp_hi", C.c_double),
        ("bead_amp_lo", C.c_double), ("bead_amp_hi", C.c_double),
        ("controller", C.c_int64),
    ]


class Sim(C.Structure):
    _fields_ = [
        ("N", C.c_int64), ("nb", C.c_int64),
        ("r", _pd), ("t3", _pd), ("t2", _pd),
        ("r_trial", _pd), ("t3_trial", _pd), ("t2_trial", _pd),
        ("states", _pl), ("states_trial", _pl), ("mods", _pl),
        ("eps_bend", _pd), ("eps_par", _pd), ("eps_perp", _pd), ("gamma", _pd), ("eta", _pd),
        ("max_binders", C.c_int64), ("mu_adjust_factor", C.c_double), ("bead_vol", C.c_double),
        ("sites_per_bead", _pl),
        ("bind_energy_mod", _pd), ("bind_energy_no_mod", _pd), ("chemical_potential", _pd),
        ("field_pref", _pd), ("e_intra", _pd), ("xpref", _pd),
        ("field_active", C.c_int64),
        ("nx", C.c_int64), ("ny", C.c_int64), ("nz", C.c_int64), ("n_bins", C.c_int64),
        ("width", C.c_double * 3), ("dxyz", C.c_double * 3),
        ("half_width", C.c_double * 3), ("half_step", C.c_double * 3),
        ("vol_bin", C.c_double),
        ("access_vol", _pd), ("density", _pd), ("density_trial", _pd),
        ("affected", _pl),
        ("confine_type", C.c_int64), ("confine_length", C.c_double), ("chi", C.c_double),
        ("vf_limit", C.c_float),
        ("touched", _pl), ("n_touched", C.c_int64), ("touch_stamp", _pl), ("stamp", C.c_int64),
        ("crng", GlibcRand), ("mt", MT19937),
        ("last_dE_poly", C.c_double), ("last_dE_field", C.c_double),
        ("last_accept", C.c_int64), ("last_u", C.c_double),
        ("eps_twist", _pd), ("twist0", _pd),
        ("philox", Philox),
        ("fast_n_points", C.c_int64),
        ("detailed", C.POINTER(Detailed)),
    ]


def build_lib(force: bool = False) -> Path:
    so = HERE / "libchromo_oracle.so"
    src = HERE / "chromo_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "-s", "-B", "libchromo_oracle.so"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build_lib()))
        ps = C.POINTER(Sim)
        pm = C.POINTER(Move)
        L.oc_srand.argtypes = [C.POINTER(GlibcRand), C.c_uint32]
        L.oc_rand.argtypes = [C.POINTER(GlibcRand)]
        L.oc_rand.restype = C.c_int32
        L.oc_mt_seed.argtypes = [C.POINTER(MT19937), C.c_uint32]
        L.oc_mt_next.argtypes = [C.POINTER(MT19937)]
        L.oc_mt_next.restype = C.c_uint32
        L.oc_mt_randint.argtypes = [C.POINTER(MT19937), C.c_int64, C.c_int64]
        L.oc_mt_randint.restype = C.c_int64
        L.oc_philox_block.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.oc_philox_init.argtypes = [ps, C.c_uint64, C.c_uint32, C.c_uint64]
        L.oc_bin_point.argtypes = [ps, _pd, _pl, _pd]
        L.oc_update_all_densities.argtypes = [ps, C.c_int]
        L.oc_field_E.argtypes = [ps]
        L.oc_field_E.restype = C.c_double
        L.oc_poly_E.argtypes = [ps]
        L.oc_poly_E.restype = C.c_double
        L.oc_field_dE.argtypes = [ps, _pl, C.c_int64, C.c_int]
        L.oc_field_dE.restype = C.c_double
        L.oc_update_affected_densities.argtypes = [ps]
        L.oc_poly_dE.argtypes = [ps, C.c_int, _pl, C.c_int64]
        L.oc_poly_dE.restype = C.c_double
        L.oc_binding_free_energy.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_double, C.c_double]
        L.oc_binding_free_energy.restype = C.c_double
        L.oc_rotation_matrix.argtypes = [_pd, _pd, C.c_double, _pd]
        L.oc_transform_rows.argtypes = [ps, _pd, _pl, C.c_int64]
        L.oc_from_point.argtypes = [C.POINTER(GlibcRand), C.c_int64, C.c_int64, C.c_int64]
        L.oc_from_point.restype = C.c_int64
        L.oc_from_left.argtypes = [C.POINTER(GlibcRand), C.c_int64, C.c_int64]
        L.oc_from_left.restype = C.c_int64
        L.oc_from_right.argtypes = [C.POINTER(GlibcRand), C.c_int64, C.c_int64]
        L.oc_from_right.restype = C.c_int64
        L.oc_propose.argtypes = [ps, C.c_int, C.c_double, C.c_int64, _pl]
        L.oc_propose.restype = C.c_int64
        L.oc_accept.argtypes = [ps, pm, C.c_int, _pl, C.c_int64]
        L.oc_reject.argtypes = [ps, pm, C.c_int, _pl, C.c_int64]
        L.oc_update_amplitudes.argtypes = [pm]
        L.oc_mc_step.argtypes = [ps, pm, C.c_int, _pl]
        L.oc_mc_step.restype = C.c_int
        L.oc_mc_sim.argtypes = [ps, pm, C.c_int64, C.c_uint32, _pl]
        L.oc_mc_sim_ordered.argtypes = [ps, pm, C.c_int64, C.c_uint32, _pl, C.POINTER(C.c_int32)]
        _LIB = L
    return _LIB


# --------------------------------------------------------------------------
# host-side derived parameters (init-time logic of the reference, restated)
# --------------------------------------------------------------------------

def dss_table() -> np.ndarray:
    """dssWLC parameter table (columns 0-5 of chromo/util/dssWLCparams)."""
    return np.load(HERE.parent / "chromo_b200" / "data" / "dsswlc_params.npy")


def bond_params(bead_length: np.ndarray, lp: float) -> dict:
    """SSWLC._find_parameters, polymers.pyx:1545-1601."""
    tab = dss_table()
    bl = np.asarray(bead_length, dtype=float)
    out = {k: np.zeros(len(bl)) for k in ("delta", "eps_bend", "gamma", "eps_par", "eps_perp", "eta")}
    for i in range(len(bl)):
        d = bl[i] / lp
        out["delta"][i] = d
        out["eps_bend"][i] = np.interp(d, tab[:, 0], tab[:, 1]) / d
        out["gamma"][i] = np.interp(d, tab[:, 0], tab[:, 2]) * d * lp
        out["eps_par"][i] = np.interp(d, tab[:, 0], tab[:, 3]) / (d * lp ** 2)
        out["eps_perp"][i] = np.interp(d, tab[:, 0], tab[:, 4]) / (d * lp ** 2)
        out["eta"][i] = np.interp(d, tab[:, 0], tab[:, 5]) / lp
    if len(bl) and np.all(bl == bl[0]):  # uniform spacing: same value everywhere
        pass
    return out


HP1 = dict(name="HP1", sites_per_bead=2, bind_energy_mod=-0.01, bind_energy_no_mod=1.52,
           interaction_energy=-4.0, chemical_potential=-1.0, interaction_radius=3.0,
           cross_talk={"PRC1": 0.0})
PRC1 = dict(name="PRC1", sites_per_bead=2, bind_energy_mod=-0.01, bind_energy_no_mod=1.52,
            interaction_energy=-4.0, chemical_potential=-1.0, interaction_radius=3.0,
            cross_talk={"HP1": 0.0})
NULL_READER = dict(name="null_reader", sites_per_bead=0, bind_energy_mod=0.0,
                   bind_energy_no_mod=0.0, interaction_energy=0.0, chemical_potential=0.0,
                   interaction_radius=0.0, cross_talk={})


def field_prefactors(binders: list, vol_bin: float):
    """UniformDensityField.init_field_energy_prefactors, fields.pyx:687-712."""
    nb = len(binders)
    pref = np.zeros(nb)
    e_intra = np.zeros(nb)
    xpref = np.zeros((nb, nb))
    for i, b in enumerate(binders):
        v_int = (4.0 / 3.0) * np.pi * b["interaction_radius"] ** 3  # binders.pyx:109
        pref[i] = 0.5 * b["interaction_energy"] * v_int * vol_bin
        e_intra[i] = b["interaction_energy"] * (1 - v_int / vol_bin)
        for j, nxt in enumerate(binders):
            if nxt["name"] in b["cross_talk"]:
                xpref[i, j] = b["cross_talk"][nxt["name"]] * v_int * vol_bin
    return pref, e_intra, xpref


def accessible_volumes(fld: dict, n_side: int = 20) -> np.ndarray:
    """UniformDensityField.get_accessible_volumes (fields.pyx:714-770) restated: with
    assume_fully_accessible = 0 and a spherical confinement, a voxel cut by the sphere
    (get_split_voxels 805-840: centre within sqrt(2)/4 of the largest voxel edge of the surface) keeps the
    fraction of an n_side^3 sub-grid -- anchored at the voxel's lower corner, define_voxel_subgrid 842-880 --
    that lies strictly inside (get_frac_accessible 882-951); every other voxel keeps vol_bin.  Pinned by the
    `access_vols` array of tests/golden/static_av.npz (the reference's own output)."""
    nx, ny, nz = fld["nx"], fld["ny"], fld["nz"]
    n_bins = nx * ny * nz
    dx, dy, dz = fld["x_width"] / nx, fld["y_width"] / ny, fld["z_width"] / nz
    vol_bin = fld["x_width"] * fld["y_width"] * fld["z_width"] / n_bins
    vols = np.full(n_bins, vol_bin)
    if fld.get("assume_fully_accessible", 1) == 1 or fld.get("confine_type", "") != "Spherical":
        return vols
    R = fld["confine_length"]
    buf = np.sqrt(2) / 4 * max(dx, dy, dz)
    k = np.arange(n_side, dtype=float)
    for b in range(n_bins):
        ix, iy, iz = b % nx, (b // nx) % ny, b // (nx * ny)
        # get_voxel_coords fields.pyx:795-799: `(nxyz[j] - 1) / 2` on C longs is FLOOR division (cdivision
        # False, language_level 2): for an even grid the "centres" sit half a voxel below the true ones
        c = np.array([(ix - (nx - 1) // 2) * dx, (iy - (ny - 1) // 2) * dy, (iz - (nz - 1) // 2) * dz])
        dist = np.sqrt(c[0] ** 2 + c[1] ** 2 + c[2] ** 2)
        if dist < R - buf or dist > R + buf:
            continue
        corner = c - np.array([dx / 2, dy / 2, dz / 2])
        gx, gy, gz = np.meshgrid(corner[0] + k * (dx / n_side), corner[1] + k * (dy / n_side),
                                 corner[2] + k * (dz / n_side), indexing="ij")
        inside = np.sqrt(gx ** 2 + gy ** 2 + gz ** 2) < R
        vols[b] = vol_bin * (inside.sum() / float(n_side ** 3))
    return vols


def nucleosome_constants(bp_wrap: float) -> dict:
    """Geometry of a DetailedNucleosome (beads.py:448-515, util/nucleo_geom.py:17-245) for one bp_wrap, restated:
    the local frame of the entering DNA, the entry / exit points on the nucleosome's super-helix and the
    coefficients that give the exiting t3 / t1 in terms of the entering (t3, t2, t1)."""
    LENGTH_BP, dens = 0.332, 10.17
    Rn = 4.1899999999999995
    h = 4.531142964071856 / 2
    s_def = (147 - 1) * LENGTH_BP
    w0 = 2 * np.pi / (dens * LENGTH_BP)
    Lt = np.sqrt(4 * np.pi ** 2 * Rn ** 2 + h ** 2)
    Phi = w0 - 2 * np.pi * h / (Lt ** 2)
    t3f = lambda s: np.array([-2 * np.pi * Rn / Lt * np.sin(2 * np.pi * s / Lt),
                              2 * np.pi * Rn / Lt * np.cos(2 * np.pi * s / Lt), h / Lt])
    nrm = lambda s: np.array([-np.cos(2 * np.pi * s / Lt), -np.sin(2 * np.pi * s / Lt), 0])
    bnm = lambda s: np.cross(t3f(s), nrm(s))
    t1f = lambda s: np.cos(Phi * s) * nrm(s) + np.sin(Phi * s) * bnm(s)
    t2f = lambda s: -np.sin(Phi * s) * nrm(s) + np.cos(Phi * s) * bnm(s)
    s = (bp_wrap - 1) * LENGTH_BP
    r_enter = np.array([Rn, 0, -(h * s_def / Lt) / 2])
    r_exit = np.array([Rn * np.cos(2 * np.pi * s / Lt), Rn * np.sin(2 * np.pi * s / Lt), h * s / Lt - ((s_def / Lt * h) / 2)])
    return dict(
        bead_rad=Rn,
        t3_local=np.array([0, 2 * np.pi * Rn / Lt, h / Lt]), t2_local=np.array([0, -h / Lt, 2 * np.pi * Rn / Lt]),
        r_enter_norm=np.linalg.norm(r_enter), r_enter_unit=r_enter / np.linalg.norm(r_enter),
        r_exit_norm=np.linalg.norm(r_exit), r_exit_unit=r_exit / np.linalg.norm(r_exit),
        a3=np.array([np.dot(t3f(s), t3f(0)), np.dot(t3f(s), t2f(0)), np.dot(t3f(s), t1f(0))]),
        a1=np.array([np.dot(t1f(s), t3f(0)), np.dot(t1f(s), t2f(0)), np.dot(t1f(s), t1f(0))]))


def amplitude_bounds(N: int, min_spacing: float):
    """get_amplitude_bounds, mc/__init__.py:295-332."""
    bead = {
        "crank_shaft": (min(30, N), min(150, N)),
        "slide": (min(10, N), min(150, N)),
        "end_pivot": (min(50, N / 4), min(150, int(N / 2))),
        "tangent_rotation": (1, N),
        "change_binding_state": (1, 1),
    }
    move = {
        "crank_shaft": (0.1 * np.pi, 0.25 * np.pi),
        "slide": (0.2 * min_spacing, 0.3 * min_spacing),
        "end_pivot": (0.2 * np.pi, 0.25 * np.pi),
        "tangent_rotation": (0.05 * np.pi, 0.2 * np.pi),
        "change_binding_state": (0, 0),
    }
    return bead, move


def make_moves(N, min_spacing, per_cycle=(30, 1, 60, 60, 10), controller=1, move_on=(1, 1, 1, 1, 1)):
    """all_moves(..., SimpleControl), mc_controller.py:216-266."""
    bead, move = amplitude_bounds(N, min_spacing)
    arr = (Move * 5)()
    for i, name in enumerate(MOVE_NAMES):
        m = arr[i]
        m.move_on = move_on[i]
        m.num_per_cycle = per_cycle[i]
        m.amp_move = move[name][0]
        m.amp_bead = int(bead[name][0])
        m.num_attempt = 0
        m.num_success = 0
        m.acceptance_rate = 0.0
        m.alpha = 2 / (20.0 + 1)
        m.move_amp_lo, m.move_amp_hi = move[name]
        m.bead_amp_lo, m.bead_amp_hi = bead[name]
        m.controller = controller
    return arr


# --------------------------------------------------------------------------
# synthetic problems
# --------------------------------------------------------------------------

def confined_walk(N, step, R, rng):
    """Vectorised-in-spirit restatement of poly_paths.confined_gaussian_walk
    (poly_paths.py:298-335): unit Gaussian-direction steps of length `step`,
    re-drawn while the bead would leave the sphere of radius R."""
    r = np.zeros((N, 3))
    if 0 < R < step:  # the first step from the origin can never stay inside: fail instead of looping for ever
        raise ValueError(f"a walk with steps of {step} does not fit a sphere of radius {R:.3g} "
                         f"(pass confine=\"\" or more beads)")
    for i in range(1, N):
        while True:
            d = rng.standard_normal(3)
            d *= step / np.linalg.norm(d)
            p = r[i - 1] + d
            if R <= 0 or np.linalg.norm(p) <= R:
                r[i] = p
                break
    return r


def tangents_from_coords(r, rng):
    """estimate_tangents_from_coordinates (poly_paths.py:506-545) in spirit:
    t3 = normalised central differences, t2 a unit vector orthogonal to t3."""
    N = len(r)
    t3 = np.zeros_like(r)
    t3[1:-1] = r[2:] - r[:-2]
    t3[0] = r[1] - r[0]
    t3[-1] = r[-1] - r[-2]
    t3 /= np.linalg.norm(t3, axis=1)[:, None]
    a = rng.standard_normal((N, 3))
    t2 = a - (a * t3).sum(1)[:, None] * t3
    t2 /= np.linalg.norm(t2, axis=1)[:, None]
    return t3, t2


def synthetic_marks(N, nb, rng, p=(0.457, 0.084, 0.459), domain=40):
    """Blocky 0/1/2 mark pattern with the marginal distribution of the
    reference's H3K9me3 track (45.7/8.4/45.9 %)."""
    mods = np.zeros((N, nb), dtype=np.int64)
    for b in range(nb):
        i = 0
        while i < N:
            L = 1 + rng.geometric(1.0 / domain)
            mods[i:i + L, b] = rng.choice(3, p=p)
            i += L
    return mods


def make_spec(N=200, nb=1, seed=0, grid=None, confine="Spherical", chi=1.0, binders=None,
              spacing=16.5, lp=53.0, bead_rad=5.0, mu=-1.2, random_states=True, vf_limit=0.5,
              max_binders=-1, cross_talk=0.0):
    """A seeded synthetic replica following SURVEY.md 8(d)."""
    rng = np.random.default_rng(seed)
    dens = 393216 / (4.0 / 3.0 * math.pi * 900.0 ** 3)
    R = (N / dens / (4.0 * math.pi / 3.0)) ** (1.0 / 3.0)
    n_acc = max(int(round(63 * R / 900.0)), 2)
    nx = n_acc + 2 if grid is None else grid
    W = 2 * R * (1 + 2.0 / n_acc)
    if binders is None:
        if nb == 1:
            binders = [dict(HP1)]
        else:
            binders = [dict(HP1), dict(PRC1)][:nb]
            binders[0]["cross_talk"] = {"PRC1": cross_talk}
        for b in binders:
            b["chemical_potential"] = mu
    r = confined_walk(N, spacing, R if confine == "Spherical" else 0.0, rng)
    t3, t2 = tangents_from_coords(r, rng)
    mods = synthetic_marks(N, nb, rng)
    if random_states:
        states = np.array([[rng.integers(0, b["sites_per_bead"] + 1) for b in binders]
                           for _ in range(N)], dtype=np.int64).reshape(N, nb)
    else:
        states = np.zeros((N, nb), dtype=np.int64)
    for j, b in enumerate(binders):
        mods[:, j] = np.minimum(mods[:, j], b["sites_per_bead"])
    return dict(
        N=N, nb=nb, r=r, t3=t3, t2=t2, states=states, mods=mods,
        bead_length=np.full(N - 1, spacing), lp=lp, bead_rad=bead_rad, binders=binders,
        max_binders=max_binders,
        field=dict(x_width=W, nx=nx, y_width=W, ny=nx, z_width=W, nz=nx,
                   confine_type=confine, confine_length=R if confine else 0.0,
                   chi=chi, vf_limit=vf_limit),
    )


# --------------------------------------------------------------------------
# the oracle object
# --------------------------------------------------------------------------

def _p(a, t):
    return a.ctypes.data_as(t)


class OracleSim:
    """One replica (polymer + field) evaluated by the C oracle."""

    def __init__(self, spec: dict, mu_adjust_factor: float = 1.0, srand_seed: int = 1):
        L = lib()
        self.L = L
        self.spec = spec
        N, nb = spec["N"], spec["nb"]
        f64 = lambda a: np.ascontiguousarray(np.array(a, dtype=np.float64))
        i64 = lambda a: np.ascontiguousarray(np.array(a, dtype=np.int64))
        self.r, self.t3, self.t2 = f64(spec["r"]), f64(spec["t3"]), f64(spec["t2"])
        self.r_trial, self.t3_trial, self.t2_trial = self.r.copy(), self.t3.copy(), self.t2.copy()
        self.states = i64(spec["states"]).reshape(N, nb)
        self.states_trial = self.states.copy()
        self.mods = i64(spec["mods"]).reshape(N, nb)
        bp = bond_params(spec["bead_length"], spec["lp"])
        self.bp = {k: f64(v) for k, v in bp.items()}
        binders = spec["binders"]
        self.sites = i64([b["sites_per_bead"] for b in binders])
        self.e_mod = f64([b["bind_energy_mod"] for b in binders])
        self.e_nomod = f64([b["bind_energy_no_mod"] for b in binders])
        self.mu = f64([b["chemical_potential"] for b in binders])
        s = Sim()
        s.N, s.nb = N, nb
        for name in ("r", "t3", "t2", "r_trial", "t3_trial", "t2_trial"):
            setattr(s, name, _p(getattr(self, name), _pd))
        s.states, s.states_trial, s.mods = (_p(self.states, _pl), _p(self.states_trial, _pl),
                                            _p(self.mods, _pl))
        for name in ("eps_bend", "eps_par", "eps_perp", "gamma", "eta"):
            setattr(s, name, _p(self.bp[name], _pd))
        if spec.get("lt") is not None:  # SSTWLC: twist modulus and natural twist per bond (polymers.pyx:2000, 2088-2090)
            bl = np.asarray(spec["bead_length"], dtype=float)
            self.eps_twist = f64(spec["lt"] / ((bl / spec["lp"]) * spec["lp"]))
            self.twist0 = f64(bl * (2 * np.pi / 10.5) / 0.332)
            s.eps_twist, s.twist0 = _p(self.eps_twist, _pd), _p(self.twist0, _pd)
        if spec.get("bp_wrap") is not None:  # DetailedChromatin (polymers.pyx:2455-2607)
            nc = nucleosome_constants(spec["bp_wrap"])
            self._detailed = Detailed()
            for k in ("t3_local", "t2_local", "r_enter_unit", "r_exit_unit", "a3", "a1"):
                getattr(self._detailed, k)[:] = list(map(float, nc[k]))
            self._detailed.r_enter_norm, self._detailed.r_exit_norm = float(nc["r_enter_norm"]), float(nc["r_exit_norm"])
            if spec.get("no_diameter"):  # DetailedChromatin2 (polymers.pyx:2627-2735): bonds run between bead
                self._detailed.r_enter_norm = self._detailed.r_exit_norm = 0.0  # centres, only the frames change
            s.detailed = C.pointer(self._detailed)
        s.max_binders = spec.get("max_binders", -1)
        s.mu_adjust_factor = mu_adjust_factor
        s.bead_vol = (4 / 3) * np.pi * spec["bead_rad"] ** 3  # beads.py:415
        s.sites_per_bead = _p(self.sites, _pl)
        s.bind_energy_mod, s.bind_energy_no_mod = _p(self.e_mod, _pd), _p(self.e_nomod, _pd)
        s.chemical_potential = _p(self.mu, _pd)
        fld = spec.get("field")
        if fld is not None:
            s.field_active = 1
            s.nx, s.ny, s.nz = fld["nx"], fld["ny"], fld["nz"]
            n_bins = s.nx * s.ny * s.nz
            s.n_bins = n_bins
            widths = [fld["x_width"], fld["y_width"], fld["z_width"]]
            ns = [fld["nx"], fld["ny"], fld["nz"]]
            for j in range(3):  # init_grid fields.pyx:536-575
                s.width[j] = widths[j]
                s.dxyz[j] = widths[j] / ns[j]
                s.half_width[j] = 0.5 * widths[j]
                s.half_step[j] = 0.5 * (widths[j] / ns[j])
            s.vol_bin = widths[0] * widths[1] * widths[2] / n_bins
            self.access_vol = np.ascontiguousarray(accessible_volumes(fld))  # uniform unless assume_fully_accessible = 0
            s.confine_type = CONFINE[fld.get("confine_type", "")]
            s.confine_length = fld.get("confine_length", 0.0)
            s.chi = fld.get("chi", 1.0)
            s.vf_limit = fld.get("vf_limit", 0.5)
            if fld.get("fast_field", 0) == 1:  # n_points is rounded up to an even number (fields.pyx:603)
                npts = int(fld.get("n_points", 1000))
                s.fast_n_points = npts + (npts % 2)
            pref, e_intra, xpref = field_prefactors(binders, s.vol_bin)
        else:
            s.field_active = 0
            n_bins = 1
            s.n_bins = 0
            self.access_vol = np.ones(1)
            pref, e_intra, xpref = np.zeros(nb), np.zeros(nb), np.zeros((nb, nb))
        self.pref, self.e_intra, self.xpref = f64(pref), f64(e_intra), f64(xpref)
        s.field_pref, s.e_intra, s.xpref = _p(self.pref, _pd), _p(self.e_intra, _pd), _p(self.xpref, _pd)
        self.density = np.zeros((n_bins, nb + 1))
        self.density_trial = np.zeros((n_bins, nb + 1))
        self.affected = np.zeros(n_bins, dtype=np.int64)
        self.touched = np.zeros(n_bins, dtype=np.int64)
        self.touch_stamp = np.zeros(n_bins, dtype=np.int64)
        s.access_vol, s.density, s.density_trial = (_p(self.access_vol, _pd), _p(self.density, _pd),
                                                    _p(self.density_trial, _pd))
        s.affected, s.touched, s.touch_stamp = (_p(self.affected, _pl), _p(self.touched, _pl),
                                                _p(self.touch_stamp, _pl))
        s.stamp = 0
        self.s = s
        self.inds = np.zeros(max(N, 1), dtype=np.int64)
        L.oc_srand(C.byref(s.crng), srand_seed)
        L.oc_mt_seed(C.byref(s.mt), 0)
        if fld is not None:  # UniformDensityField.__init__ fields.pyx:532
            L.oc_update_all_densities(C.byref(s), 1)

    # -- RNG
    def srand(self, seed):
        self.L.oc_srand(C.byref(self.s.crng), seed)

    def np_seed(self, seed):
        self.L.oc_mt_seed(C.byref(self.s.mt), seed)

    def use_production_streams(self, seed, replica=0, next_attempt=0):
        """Draw from the product's production streams (Philox4x32-10 keyed by seed and replica, one
        stream per attempt; chromo_b200/csrc/rng.cuh) instead of the reference's rand() / MT19937, so
        that a production-mode run of the product can be replayed attempt for attempt.  `srand`
        switches back."""
        self.L.oc_philox_init(C.byref(self.s), seed, replica, next_attempt)

    @property
    def attempts_made(self):
        return int(self.s.philox.next_attempt)

    # -- A1
    def bin_point(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        idx = np.zeros(8, dtype=np.int64)
        w = np.zeros(8)
        self.L.oc_bin_point(C.byref(self.s), _p(xyz, _pd), _p(idx, _pl), _p(w, _pd))
        return idx, w

    # -- A8
    def update_all_densities(self, for_all_polymers=False):
        self.L.oc_update_all_densities(C.byref(self.s), int(for_all_polymers))

    def field_E(self):
        return self.L.oc_field_E(C.byref(self.s))

    def poly_E(self):
        return self.L.oc_poly_E(C.byref(self.s))

    # -- A1-A7, A9, A10
    def field_dE(self, inds, state_change):
        inds = np.ascontiguousarray(inds, dtype=np.int64)
        dE = self.L.oc_field_dE(C.byref(self.s), _p(inds, _pl), len(inds), int(state_change))
        return dE, self.touched[: self.s.n_touched].copy()

    def poly_dE(self, move, inds):
        inds = np.ascontiguousarray(inds, dtype=np.int64)
        return self.L.oc_poly_dE(C.byref(self.s), move, _p(inds, _pl), len(inds))

    def commit_field(self):
        self.L.oc_update_affected_densities(C.byref(self.s))

    # -- A11 / A12
    def propose(self, move, amp_move, amp_bead):
        n = self.L.oc_propose(C.byref(self.s), move, amp_move, amp_bead, _p(self.inds, _pl))
        return self.inds[:n].copy()

    def mc_step(self, mv, move):
        return self.L.oc_mc_step(C.byref(self.s), C.byref(mv), move, _p(self.inds, _pl))

    def mc_sim(self, moves, num_mc_steps, np_seed, order=None):
        """`order`: move ids in the order of the reference's controller list (default: the canonical one)."""
        if order is None:
            self.L.oc_mc_sim(C.byref(self.s), moves, num_mc_steps, np_seed, _p(self.inds, _pl))
        else:
            o = (C.c_int32 * 5)(*[int(x) for x in order])
            self.L.oc_mc_sim_ordered(C.byref(self.s), moves, num_mc_steps, np_seed, _p(self.inds, _pl), o)


# --------------------------------------------------------------------------
# the reference's own objects for the same spec (authoring container / any
# box where oracle/_ref has been built)
# --------------------------------------------------------------------------

def ref_available() -> bool:
    return (HERE / "_ref" / ".built").exists()


def ref_objects(spec: dict):
    """Build reference `Chromatin`, binder collection and
    `UniformDensityField`/`NullField` for `spec`.  Returns (poly, binders_df,
    field, modules)."""
    import sys
    import importlib
    sys.path.insert(0, str(HERE))
    import build_ref
    build_ref.activate()
    import chromo.polymers as ply
    import chromo.binders as bnd
    import chromo.fields as fld
    objs = []
    for b in spec["binders"]:
        o = bnd.get_by_name(b["name"])
        for k in ("sites_per_bead", "bind_energy_mod", "bind_energy_no_mod", "interaction_energy",
                  "chemical_potential", "interaction_radius"):
            setattr(o, k, b[k])
        o.interaction_volume = (4.0 / 3.0) * np.pi * b["interaction_radius"] ** 3
        o.cross_talk_interaction_energy = dict(b["cross_talk"])
        o.cross_talk_field_energy_prefactor = {}
        objs.append(o)
    df = bnd.make_binder_collection(objs)
    N, nb = spec["N"], spec["nb"]
    kw = dict(
        bead_length=np.ascontiguousarray(spec["bead_length"], dtype=float),
        bead_rad=float(spec["bead_rad"]),
        t3=np.ascontiguousarray(spec["t3"], dtype=float).copy(),
        t2=np.ascontiguousarray(spec["t2"], dtype=float).copy(),
        states=np.ascontiguousarray(spec["states"], dtype=np.int64).reshape(N, nb).copy(),
        binder_names=np.array([b["name"] for b in spec["binders"]]),
        chemical_mods=np.ascontiguousarray(spec["mods"], dtype=np.int64).reshape(N, nb).copy(),
        chemical_mod_names=np.array([f"mod{j}" for j in range(nb)]),
        max_binders=spec.get("max_binders", -1),
    )
    r = np.ascontiguousarray(spec["r"], dtype=float).copy()
    if spec.get("bp_wrap") is not None:
        kw.pop("bead_rad")  # fixed by the nucleosome geometry (consts_dict["R"])
        cls = ply.DetailedChromatin2 if spec.get("no_diameter") else ply.DetailedChromatin
        poly = cls("replica", r, bp_wrap=float(spec["bp_wrap"]), lp=float(spec["lp"]), lt=float(spec["lt"]), **kw)
    elif spec.get("lt") is not None:
        poly = ply.SSTWLC("replica", r, lp=float(spec["lp"]), lt=float(spec["lt"]), **kw)
    elif spec["lp"] == 53.0:
        poly = ply.Chromatin("replica", r, **kw)
    else:
        poly = ply.SSWLC("replica", r, lp=float(spec["lp"]), **kw)
    f = spec.get("field")
    if f is None:
        field = fld.NullField()
    else:
        field = fld.UniformDensityField(
            [poly], df, f["x_width"], f["nx"], f["y_width"], f["ny"], f["z_width"], f["nz"],
            confine_type=f.get("confine_type", ""), confine_length=f.get("confine_length", 0.0),
            chi=f.get("chi", 1.0), vf_limit=f.get("vf_limit", 0.5),
            assume_fully_accessible=f.get("assume_fully_accessible", 1),
            fast_field=f.get("fast_field", 0), n_points=f.get("n_points", 1000))
    mods = dict(polymers=ply, binders=bnd, fields=fld,
                mc_sim=importlib.import_module("chromo.mc.mc_sim"),
                mc_controller=importlib.import_module("chromo.mc.mc_controller"),
                mc=importlib.import_module("chromo.mc"),
                move_funcs=importlib.import_module("chromo.mc.move_funcs"),
                shim=importlib.import_module("oracle_shim"))
    return poly, df, field, mods
