"""The five Monte-Carlo moves as named callables (chromo/mc/move_funcs.pyx:
crank_shaft 39, end_pivot 285, slide 403, tangent_rotation 470,
change_binding_state 717).

Calling one proposes a move for `polymer` ON THE GPU (the proposal code lives in
chromo_b200/csrc/mc_kernel.cuh) and, like the reference, leaves the trial state
in `polymer.r_trial / t3_trial / t2_trial / states_trial` and returns the moved
bead indices.  The hot path never goes through these wrappers: `mc_sim` runs
proposal, energies and Metropolis fused in one kernel.
"""
import numpy as np

from .._lib import MOVE_ID, RNG_REPLAY


def _propose(name, polymer, amp_move, amp_bead):
    field = getattr(polymer, "_field", None)
    if field is None:
        from ..fields import NullField
        field = NullField([polymer])
        polymer._field = field
    e = field._push(polymer)
    from .mc_sim import _seed_for, rng_mode
    out = e.mc_step(0, MOVE_ID[name], amp_move, amp_bead, float(polymer.mu_adjust_factor),
                    rng_mode(), _seed_for(polymer), force_accept=0)
    inds = out["inds"]
    rows = out["rows"]
    polymer.r_trial[inds] = rows[:, 0:3]
    polymer.t3_trial[inds] = rows[:, 3:6]
    polymer.t2_trial[inds] = rows[:, 6:9]
    polymer.states_trial[inds] = rows[:, 9:].astype(np.int64)
    polymer.last_amp_bead = len(inds)
    polymer._last_step = out
    return inds


def crank_shaft(polymer, amp_move, amp_bead):
    """Rotate a segment about the axis through its end beads."""
    return _propose("crank_shaft", polymer, amp_move, amp_bead)


def end_pivot(polymer, amp_move, amp_bead):
    """Rotate a segment at one end of the chain about a random axis."""
    return _propose("end_pivot", polymer, amp_move, amp_bead)


def slide(polymer, amp_move, amp_bead):
    """Translate a segment in a random direction."""
    return _propose("slide", polymer, amp_move, amp_bead)


def tangent_rotation(polymer, amp_move, amp_bead):
    """Rotate the tangents of randomly selected beads."""
    return _propose("tangent_rotation", polymer, amp_move, amp_bead)


def change_binding_state(polymer, amp_move, amp_bead):
    """Redraw the reader-protein binding state of a segment."""
    return _propose("change_binding_state", polymer, amp_move, amp_bead)


def transform_r_t3_t2(polymer, inds, n_inds):
    """Apply `polymer.transformation_mat` to the listed beads' r, t3, t2 and
    store the result in the trial arrays (move_funcs.pyx:121-154): a
    deterministic host-side helper kept for API parity (the kernels apply the
    same affine map per lane)."""
    M = np.asarray(polymer.transformation_mat)
    for i in range(n_inds):
        b = inds[i]
        for j in range(3):
            er = e3 = e2 = 0.0
            for k in range(3):
                er += M[j, k] * polymer.r[b, k]
                e3 += M[j, k] * polymer.t3[b, k]
                e2 += M[j, k] * polymer.t2[b, k]
            polymer.r_trial[b, j] = er + M[j, 3]
            polymer.t3_trial[b, j] = e3
            polymer.t2_trial[b, j] = e2
