"""Helpers shared by the GPU parity tests, the emulator tests, smoke() and bench."""
import numpy as np

import oracle as O
from chromo_b200._lib import MOVE_DTYPE
from chromo_b200.engine import Engine

_MV_FIELDS = ("amp_move", "move_amp_lo", "move_amp_hi", "bead_amp_lo", "bead_amp_hi", "acceptance_rate",
              "alpha", "num_attempt", "num_success", "amp_bead", "num_per_cycle", "move_on", "controller")


def engine_from_spec(spec, R=1, device=0, chi=None, mu=None):
    """R identical replicas of `spec` on one device, densities initialised as
    UniformDensityField.__init__ does (fields.pyx:532)."""
    N, nb = spec["N"], spec["nb"]
    f = spec["field"]
    bead_vol = (4 / 3) * np.pi * spec["bead_rad"] ** 3
    e = Engine(R, N, nb, grid=f, bead_vol=bead_vol, max_binders=spec.get("max_binders", -1), device=device)
    if f is not None and f["nx"] * f["ny"] * f["nz"] > 0:
        vol_bin = f["x_width"] * f["y_width"] * f["z_width"] / (f["nx"] * f["ny"] * f["nz"])
    else:  # NullField (nx = ny = nz = 0 keeps the confinement, fields.pyx:280-318)
        vol_bin = 1.0
    pref, e_intra, xpref = O.field_prefactors(spec["binders"], vol_bin)
    e.set_binders(spec["binders"], pref, e_intra, xpref)
    bp = O.bond_params(spec["bead_length"], spec["lp"])
    e.set_bond_params(bp["eps_bend"], bp["eps_par"], bp["eps_perp"], bp["gamma"], bp["eta"])
    if spec.get("lt") is not None:  # SSTWLC (polymers.pyx:2000, 2088-2090)
        bl = np.asarray(spec["bead_length"], dtype=float)
        e.set_twist_params(spec["lt"] / ((bl / spec["lp"]) * spec["lp"]), bl * (2 * np.pi / 10.5) / 0.332)
    if spec.get("bp_wrap") is not None:  # DetailedChromatin (polymers.pyx:2455-2607)
        from chromo_b200.util.nucleo_geom import nucleosome_constants
        e.set_detailed_nucleosomes(nucleosome_constants(spec["bp_wrap"], not spec.get("no_diameter")))
    e.set_replica_params(chi=(f["chi"] if f is not None else 1.0) if chi is None else chi,
                         mu=[b["chemical_potential"] for b in spec["binders"]] if mu is None else mu)
    if f is not None and f.get("assume_fully_accessible", 1) == 0:  # per-voxel accessible volumes (fields.pyx:714-951)
        from chromo_b200.fields import accessible_volumes
        e.set_access_volumes(accessible_volumes(f, 20, 0))
    if f is not None and f.get("fast_field", 0) == 1:  # sub-bin quantised binning (fields.pyx:577-671)
        e.set_fast_field(f.get("n_points", 1000))
    tile = lambda a: np.broadcast_to(np.asarray(a), (R,) + np.asarray(a).shape).copy()
    e.upload(tile(spec["r"]), tile(spec["t3"]), tile(spec["t2"]), tile(spec["states"]), tile(spec["mods"]))
    if f is not None and f["nx"] * f["ny"] * f["nz"] > 0:
        e.field_recompute(clamp=True)
    return e


def moves_array(spec, R, per_cycle=(30, 1, 60, 60, 10), controller=1, move_on=(1, 1, 1, 1, 1)):
    mv = O.make_moves(spec["N"], float(np.min(spec["bead_length"])), per_cycle=per_cycle,
                      controller=controller, move_on=move_on)
    a = np.zeros((R, 5), dtype=MOVE_DTYPE)
    for i in range(5):
        for f in _MV_FIELDS:
            a[f][:, i] = getattr(mv[i], f)
    return a
