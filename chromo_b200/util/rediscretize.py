"""Coarse-grain / refine pipeline (chromo/util/rediscretize.py), the step either
side of the MC path for chromosome-scale runs.

Same function names, arguments and results as the reference; the arithmetic
runs in the CUDA kernels of `csrc/rediscretize.cu` through the C ABI
(`chromo_cg_chromatin`, `chromo_refine_path`,
`chromo_enforce_spherical_confinement`), which are batched over replicas: the
`*_ensemble` functions at the bottom coarse-grain / refine every replica of a
`ReplicaEnsemble` in one launch, the reference-named functions are the R = 1
case.  There is no CPU path: without the CUDA library every call raises.

Randomness: like the reference, `get_refined_path` consumes numpy's global
legacy generator (`np.random.standard_normal`) in the reference's draw order, so a
script that calls `np.random.seed(s)` first gets the reference's path; pass
`seed=` to draw on the device instead (Philox4x32-10 + Box-Muller).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd

from .. import _lib
from .._lib import check, dptr, lptr


# ---- intervals (host metadata only) ------------------------------------------
def get_cg_bead_intervals(num_beads: int, cg_factor: float) -> Dict[int, Tuple[int, int]]:
    """rediscretize.py:24-54."""
    num_intervals = int(np.floor(num_beads / cg_factor))
    left_over = num_beads - num_intervals * cg_factor
    out = {i: (i * cg_factor, (i + 1) * cg_factor) for i in range(num_intervals)}
    if left_over > 0:
        end = out[num_intervals - 1][1]
        out[num_intervals] = (end, end + left_over + 1)
    return out


def _regular_factor(intervals, num_rows):
    """The device kernels take (num_rows, cg_factor); recover it from an
    interval table and insist that the table is the one
    `get_cg_bead_intervals` builds."""
    k = intervals[0][1] - intervals[0][0]
    if int(k) != k or k < 1:
        raise NotImplementedError("intervals must have integer bounds (numpy slicing in the reference needs them too)")
    k = int(k)
    if dict(intervals) != get_cg_bead_intervals(num_rows, k):
        raise NotImplementedError("only the interval tables of get_cg_bead_intervals are supported on the device")
    return k


def _cg(r, t3, states, mods, k, r_div, device=0):
    """[R,N,.] arrays -> coarse-grained [R,M,.] arrays, one launch."""
    L = _lib.lib()
    r = np.ascontiguousarray(r, dtype=np.float64)
    t3 = np.ascontiguousarray(t3, dtype=np.float64)
    R, N = r.shape[0], r.shape[1]
    M = L.chromo_cg_num_beads(N, k)
    if M < 1:
        raise ValueError(f"Coarse-graining factor {k} is not supported by a polymer of {N} beads.")
    nb = 0
    for a in (states, mods):
        if a is not None:
            nb = a.shape[2]
    st = None if states is None or nb == 0 else np.ascontiguousarray(states, dtype=np.int64)
    md = None if mods is None or nb == 0 else np.ascontiguousarray(mods, dtype=np.int64)
    r_cg, t3_cg, t2_cg = (np.empty((R, M, 3)) for _ in range(3))
    st_cg = None if st is None else np.empty((R, M, nb), dtype=np.int64)
    md_cg = None if md is None else np.empty((R, M, nb), dtype=np.int64)
    ms = np.zeros(1)
    check(L.chromo_cg_chromatin(device, R, N, nb, k, float(r_div), dptr(r), dptr(t3), lptr(st), lptr(md),
                                dptr(r_cg), dptr(t3_cg), dptr(t2_cg), lptr(st_cg), lptr(md_cg), dptr(ms)))
    return r_cg, t3_cg, t2_cg, st_cg, md_cg, float(ms[0])


def get_avg_in_intervals(r: np.ndarray, intervals) -> np.ndarray:
    """rediscretize.py:57-84."""
    k = _regular_factor(intervals, len(r))
    return _cg(r[None], r[None], None, None, k, 1.0)[0][0]


def get_orientations_in_intervals(t3: np.ndarray, intervals) -> Tuple[np.ndarray, np.ndarray]:
    """rediscretize.py:87-122."""
    k = _regular_factor(intervals, len(t3))
    out = _cg(t3[None], t3[None], None, None, k, 1.0)
    return out[1][0], out[2][0]


def get_majority_state_in_interval(states: np.ndarray, intervals) -> np.ndarray:
    """rediscretize.py:125-160."""
    k = _regular_factor(intervals, len(states))
    dummy = np.ones((1, len(states), 3))
    return _cg(dummy, dummy, np.asarray(states)[None], None, k, 1.0)[3][0]


def get_random_state_in_interval(states: np.ndarray, intervals) -> np.ndarray:
    """rediscretize.py:163-196: one `np.random.choice` per interval and column
    (host RNG bookkeeping; no arithmetic)."""
    out = np.zeros((len(intervals), states.shape[1]), dtype=int)
    for ind, (lo, hi) in intervals.items():
        seg = states[lo:hi]
        for i in range(states.shape[1]):
            out[ind, i] = np.random.choice(seg[:, i]) if hi - lo > 1 else seg[0, i]
    return out


# ---- binders and fields --------------------------------------------------------
def _copy_binders(df: pd.DataFrame) -> pd.DataFrame:
    out = df.copy()
    out["interaction_energy"] *= 1
    out["interaction_radius"] *= 1
    out["interaction_volume"] *= 1
    return out


def get_cg_binders(binders_refined: pd.DataFrame, cg_factor: float) -> pd.DataFrame:
    """rediscretize.py:364-398 (all scalings are x1 in the reference)."""
    return _copy_binders(binders_refined)


def refine_binders(binders_cg: pd.DataFrame, cg_factor: float) -> pd.DataFrame:
    """rediscretize.py:958-991."""
    return _copy_binders(binders_cg)


def cg_grid(d: dict, cg_factor: float) -> dict:
    """Grid of the coarse-grained field (get_cg_udf rediscretize.py:230-250)."""
    f = cg_factor ** (1 / 3)
    n = [int(round(d[k] / f)) for k in ("nx", "ny", "nz")]
    if min(n) < 2:
        raise ValueError(f"Coarse-graining factor {cg_factor} is not supported by the field.")
    if d["confine_type"] != "":
        n = [v + 1 for v in n]
    return dict(nx=n[0], ny=n[1], nz=n[2],
                x_width=n[0] * d["x_width"] / d["nx"], y_width=n[1] * d["y_width"] / d["ny"],
                z_width=n[2] * d["z_width"] / d["nz"],
                confine_type=d["confine_type"], confine_length=d["confine_length"] / f)


def refined_grid(d: dict, cg_factor: float) -> dict:
    """Grid of the refined field (refine_udf rediscretize.py:870-883); `d` has
    nx, ny, nz, dx, dy, dz, confine_type, confine_length."""
    f = cg_factor ** (1 / 3)
    n = [int(round(d[k] * f)) for k in ("nx", "ny", "nz")]
    if d["confine_type"] != "":
        n = [v + 1 for v in n]
    return dict(nx=n[0], ny=n[1], nz=n[2], x_width=n[0] * d["dx"], y_width=n[1] * d["dy"], z_width=n[2] * d["dz"],
                confine_type=d["confine_type"], confine_length=d["confine_length"] * f)


def get_cg_udf(udf_refined_dict: Dict, binders_refined: pd.DataFrame, cg_factor: float, polymers_cg: List):
    """rediscretize.py:199-272."""
    from ..fields import UniformDensityField
    g = cg_grid(udf_refined_dict, cg_factor)
    d = udf_refined_dict
    return UniformDensityField(
        polymers=polymers_cg, binders=get_cg_binders(binders_refined, cg_factor),
        x_width=g["x_width"], nx=g["nx"], y_width=g["y_width"], ny=g["ny"], z_width=g["z_width"], nz=g["nz"],
        confine_type=g["confine_type"], confine_length=g["confine_length"], chi=d["chi"],
        assume_fully_accessible=d["assume_fully_accessible"], vf_limit=d["vf_limit"],
        fast_field=d["fast_field"], n_points=d["n_points"])


def refine_udf(udf_cg, binders_cg: pd.DataFrame, cg_factor: float, polymers_refined: List):
    """rediscretize.py:845-900."""
    from ..fields import UniformDensityField
    g = refined_grid(dict(nx=udf_cg.nx, ny=udf_cg.ny, nz=udf_cg.nz, dx=udf_cg.dx, dy=udf_cg.dy, dz=udf_cg.dz,
                          confine_type=udf_cg.confine_type, confine_length=udf_cg.confine_length), cg_factor)
    return UniformDensityField(
        polymers=polymers_refined, binders=refine_binders(binders_cg, cg_factor),
        x_width=g["x_width"], nx=g["nx"], y_width=g["y_width"], ny=g["ny"], z_width=g["z_width"], nz=g["nz"],
        confine_type=g["confine_type"], confine_length=g["confine_length"], chi=udf_cg.chi,
        assume_fully_accessible=udf_cg.assume_fully_accessible, vf_limit=udf_cg.vf_limit,
        fast_field=udf_cg.fast_field, n_points=udf_cg.n_points)


# ---- coarse-graining of one polymer -----------------------------------------------
def get_cg_chromatin(polymer, cg_factor: float, name_cg: Optional[str] = "Chr_CG",
                     random_states: Optional[bool] = False):
    """rediscretize.py:401-471."""
    from ..polymers import Chromatin
    if int(cg_factor) != cg_factor:
        raise TypeError("slice indices must be integers (the reference slices with i * cg_factor)")
    k = int(cg_factor)
    intervals = get_cg_bead_intervals(polymer.num_beads, k)
    mods = None if random_states else np.asarray(polymer.chemical_mods)[None]
    r_cg, t3_cg, t2_cg, st_cg, md_cg, _ = _cg(np.asarray(polymer.r)[None], np.asarray(polymer.t3)[None],
                                              np.asarray(polymer.states)[None], mods, k, cg_factor ** (1 / 3))
    chem = get_random_state_in_interval(np.asarray(polymer.chemical_mods), intervals) if random_states else md_cg[0]
    return Chromatin(
        name=name_cg, r=r_cg[0], bead_length=np.ones(r_cg.shape[1] - 1) * polymer.bead_length[0],
        bead_rad=polymer.bead_rad, t3=t3_cg[0], t2=t2_cg[0], states=st_cg[0],
        binder_names=polymer.binder_names, chemical_mods=np.ascontiguousarray(chem, dtype=np.int64),
        chemical_mod_names=polymer.chemical_mod_names, log_path=polymer.log_path,
        max_binders=polymer.max_binders)


# ---- refinement ---------------------------------------------------------------------
def get_refined_intervals(cg_r: np.ndarray, num_beads_cg: int, num_beads_refined: int):
    """rediscretize.py:537-583 (host metadata: steps per segment and end points)."""
    seg = int(np.floor(num_beads_refined / (num_beads_cg - 1)))
    h1 = int(np.floor(seg / 2))
    h2 = seg - h1
    left = num_beads_refined % (num_beads_cg - 1)
    num_steps = {0: h1}
    num_steps.update({i: seg for i in range(1, num_beads_cg - 1)})
    num_steps[num_beads_cg - 1] = h2 - 1
    nan = np.asarray([np.nan] * 3)
    start_end = {0: (nan, cg_r[0, :])}
    start_end.update({i: (cg_r[i - 1, :], cg_r[i, :]) for i in range(1, num_beads_cg)})
    if left > 0:
        num_steps[num_beads_cg] = left
        start_end[num_beads_cg] = (cg_r[num_beads_cg - 1, :], nan)
    return num_steps, start_end


def refined_num_points(num_beads_cg: int, num_beads_refined: int) -> int:
    """Rows `get_refined_path` returns: `num_beads_refined`, or one fewer when it
    is a multiple of `num_beads_cg - 1` (the last bridge gets half a segment
    minus one, rediscretize.py:571)."""
    p = _lib.lib().chromo_refined_num_points(num_beads_cg, num_beads_refined)
    if p < 0:
        raise ZeroDivisionError("float division by zero (brownian_bridge with zero steps: fewer than 3 refined "
                                "beads per coarse-grained bond)")
    return int(p)


def _refine(cg, num_beads_refined, spacing, *, orientations, out_scale=1.0, seed=None, device=0):
    """[R,M,3] coarse path -> [R,P,3] refined path (and t2 when `orientations`)."""
    L = _lib.lib()
    cg = np.ascontiguousarray(cg, dtype=np.float64)
    R, M = cg.shape[0], cg.shape[1]
    P = refined_num_points(M, num_beads_refined)
    D = int(L.chromo_refined_num_draws(M, num_beads_refined))
    xi = None
    if seed is None:  # the reference's generator and draw order (one flat stream per path)
        xi = np.ascontiguousarray(np.random.standard_normal((R, D, 3)))
    out = np.empty((R, P, 3))
    out_t2 = np.empty((R, P, 3)) if orientations else None
    ms = np.zeros(1)
    check(L.chromo_refine_path(device, R, M, num_beads_refined, float(spacing), dptr(cg), dptr(xi),
                               0 if seed is None else int(seed) & (2 ** 64 - 1), float(out_scale),
                               1 if orientations else 0, dptr(out), dptr(out_t2), dptr(ms)))
    return out, out_t2, float(ms[0])


def get_refined_path(cg_r: np.ndarray, num_beads_refined: int, bead_spacing: Optional[float] = np.pi,
                     seed: Optional[int] = None) -> np.ndarray:
    """rediscretize.py:756-807."""
    return _refine(np.asarray(cg_r)[None], num_beads_refined, bead_spacing, orientations=False, seed=seed)[0][0]


def get_refined_orientations(t3_cg: np.ndarray, num_beads_refined: int, seed: Optional[int] = None):
    """rediscretize.py:810-842."""
    t3, t2, _ = _refine(np.asarray(t3_cg)[None], num_beads_refined, np.pi, orientations=True, seed=seed)
    return t3[0], t2[0]


def brownian_bridge(N: int, p0: np.ndarray, p1: np.ndarray, avg_step_target: Optional[float] = None) -> np.ndarray:
    """rediscretize.py:586-685: N steps from p0 to p1 (N + 1 points), average step normalised to
    `avg_step_target`.  Runs as the bridge segment of a two-bead path on the device and consumes
    numpy's generator like the reference (N - 1 Gaussian triples)."""
    p0, p1 = np.asarray(p0, dtype=float), np.asarray(p1, dtype=float)
    if N == 1:
        return np.array([p0, p1])
    if avg_step_target is None:
        raise NotImplementedError("brownian_bridge without a step target is not used by the pipeline")
    L = _lib.lib()
    n_ref = 2 * N + 2  # two coarse beads: free end of N + 1 steps, then one bridge of N steps
    h1 = N + 1
    xi = np.ones((1, h1 + N - 1, 3))
    xi[0, h1:] = np.random.standard_normal((N - 1, 3))
    out = np.empty((1, int(L.chromo_refined_num_points(2, n_ref)), 3))
    check(L.chromo_refine_path(0, 1, 2, n_ref, float(avg_step_target), dptr(np.ascontiguousarray([[p0, p1]])),
                               dptr(xi), 0, 1.0, 0, dptr(out), None, None))
    return np.vstack([out[0, h1:h1 + N], p1])


def gaussian_walk_from_point(start, N, step_size):
    """rediscretize.py:688-705: N unit-direction Gaussian steps of length `step_size` away from `start`
    (N + 1 points), as the free end of a two-bead path on the device."""
    step = np.unique(np.asarray(step_size, dtype=float))
    if len(step) != 1:
        raise NotImplementedError("one step length per walk")
    if N == 0:
        return np.asarray(start, dtype=float)[None].copy()
    L = _lib.lib()
    start = np.asarray(start, dtype=float)
    n_ref = 2 * N + 6  # free end of N + 3 steps; the walk is its first N
    h1 = N + 3
    D = int(L.chromo_refined_num_draws(2, n_ref))
    xi = np.ones((1, D, 3))
    xi[0, :N] = np.random.standard_normal((N, 3))
    out = np.empty((1, int(L.chromo_refined_num_points(2, n_ref)), 3))
    check(L.chromo_refine_path(0, 1, 2, n_ref, float(step[0]), dptr(np.ascontiguousarray([[start, start + 1.0]])),
                               dptr(xi), 0, 1.0, 0, dptr(out), None, None))
    # the free end is emitted end-first: rows h1-1 ... 0 hold points 1 ... h1
    return np.vstack([start, out[0, :h1][::-1][:N]])


def enforce_spherical_confinement(r: np.ndarray, rad: float) -> np.ndarray:
    """rediscretize.py:708-753 (in place, returns `r`)."""
    buf = np.ascontiguousarray(r, dtype=np.float64)
    batch = buf if buf.ndim == 3 else buf[None]
    check(_lib.lib().chromo_enforce_spherical_confinement(0, batch.shape[0], batch.shape[1], dptr(batch),
                                                          float(rad), None))
    if buf is not r:
        r[...] = buf
    return r


def refine_chromatin(polymer_cg, num_beads_refined: int, bead_spacing: float, chemical_mods: np.ndarray, udf_cg,
                     binding_equilibration: Optional[int] = 0, name_refine: Optional[str] = "Chr",
                     output_dir: Optional[str] = ".", seed: Optional[int] = None):
    """rediscretize.py:994-1107.  The binding equilibration (the reference's loop
    of `mc_step` calls with a binding-only move, 1093-1106) is ONE launch of
    the MC kernel with `binding_equilibration` binding-state attempts."""
    from ..polymers import Chromatin
    from ..mc import mc_controller as ctrl
    from ..mc import move_funcs as mv
    from ..mc.mc_sim import mc_sim
    num_beads_cg = len(polymer_cg.r)
    scaling = (num_beads_refined / num_beads_cg) ** (1 / 3)
    spacing_in = (np.ones(num_beads_cg - 1) * bead_spacing / scaling)[0]
    r_ref, _, _ = _refine(np.asarray(polymer_cg.r)[None], num_beads_refined, spacing_in, orientations=False,
                          out_scale=scaling, seed=seed)
    t3_ref, t2_ref, _ = _refine(np.asarray(polymer_cg.t3)[None], num_beads_refined, np.pi, orientations=True,
                                seed=None if seed is None else seed + 1)
    r_ref = r_ref[0]
    chemical_mods = np.ascontiguousarray(chemical_mods, dtype=np.int64)
    polymer = Chromatin(
        name=name_refine, r=r_ref, bead_length=np.ones(len(r_ref) - 1) * bead_spacing,
        bead_rad=polymer_cg.bead_rad, t3=t3_ref[0], t2=t2_ref[0],
        states=np.zeros((num_beads_refined, chemical_mods.shape[1]), dtype=np.int64),
        binder_names=polymer_cg.binder_names, chemical_mods=chemical_mods,
        chemical_mod_names=polymer_cg.chemical_mod_names, max_binders=polymer_cg.max_binders)
    udf = refine_udf(udf_cg, udf_cg.binders, num_beads_refined / num_beads_cg, [polymer])
    if udf.confine_type == "Spherical":
        # like the reference, the field keeps the densities of the unconfined path (1086-1089)
        polymer.r = enforce_spherical_confinement(np.asarray(polymer.r), udf.confine_length)
    if binding_equilibration > 0:
        binding_move = ctrl.specific_move(
            mv.change_binding_state, log_dir=output_dir, bead_amp_bounds={"change_binding_state": (1, 1)},
            move_amp_bounds={"change_binding_state": (1, 1)}, controller=ctrl.NoControl)
        binding_move[0].move.num_per_cycle = int(binding_equilibration)
        mc_sim([polymer], udf.binders, 1, binding_move, udf, getattr(polymer, "mu_adjust_factor", 1.0),
               0 if seed is None else seed)
    return polymer, udf


# ---- batched over replicas ------------------------------------------------------------
def coarse_grain_ensemble(r, t3, states, chemical_mods, cg_factor: int, device: int = 0):
    """`get_cg_chromatin` for [R,N,.] replica batches in one launch.  Returns a
    dict with r, t3, t2, states, chemical_mods of the coarse-grained replicas and
    `kernel_ms`, the device time of the kernel."""
    r_cg, t3_cg, t2_cg, st, md, ms = _cg(r, t3, states, chemical_mods, int(cg_factor), cg_factor ** (1 / 3), device)
    return dict(r=r_cg, t3=t3_cg, t2=t2_cg, states=st, chemical_mods=md, kernel_ms=ms)


def refine_ensemble(r_cg, t3_cg, num_beads_refined: int, bead_spacing: float, confine_length: float = 0.0,
                    seed: Optional[int] = 0, device: int = 0):
    """The geometric part of `refine_chromatin` for [R,M,3] replica batches: refined
    positions (scaled outwards, optionally pulled back into a sphere of radius
    `confine_length`) and orientations; device-side Philox deviates unless
    `seed is None` (then numpy's global generator, replica after replica)."""
    M = r_cg.shape[1]
    scaling = (num_beads_refined / M) ** (1 / 3)
    r, _, ms1 = _refine(r_cg, num_beads_refined, bead_spacing / scaling, orientations=False, out_scale=scaling,
                        seed=seed, device=device)
    t3, t2, ms2 = _refine(t3_cg, num_beads_refined, np.pi, orientations=True,
                          seed=None if seed is None else seed + 1, device=device)
    ms3 = 0.0
    if confine_length > 0:
        t = np.zeros(1)
        check(_lib.lib().chromo_enforce_spherical_confinement(device, r.shape[0], r.shape[1], dptr(r),
                                                              float(confine_length), dptr(t)))
        ms3 = float(t[0])
    return dict(r=r, t3=t3, t2=t2, kernel_ms=ms1 + ms2 + ms3)
