"""Drop-in `mc_sim` / `mc_step` (chromo/mc/mc_sim.pyx:26-182).

Same signatures, same in-place effects: after the call the polymers' `r, t3,
t2, states` (+ `*_trial`), the field's `density`, and every controller's
`move.num_attempt / num_success / amp_move / amp_bead` and
`acceptance_tracker.acceptance_rate` hold what the reference would have written.
The work itself is ONE launch of the fused CUDA kernel per call.

For many replicas use `chromo_b200.ensemble.ReplicaEnsemble`, which keeps the
state resident in HBM between calls; this single-polymer entry point uploads
and downloads every call (host arrays in, host arrays out), exactly like
handing the reference its numpy-backed memoryviews.
"""
import numpy as np

from .._lib import MOVE_DTYPE, MOVE_ID, MOVE_NAMES, NUM_MOVES, RNG_PHILOX, RNG_REPLAY

_RNG = {"mode": RNG_PHILOX}


def set_rng_mode(mode: str):
    """'philox' (default, production) or 'replay' (the reference's glibc rand() +
    numpy MT19937 streams, for draw-for-draw comparison with the reference)."""
    _RNG["mode"] = {"philox": RNG_PHILOX, "replay": RNG_REPLAY}[mode]


def rng_mode():
    return _RNG["mode"]


def _seed_for(poly):
    return int(getattr(poly, "_philox_seed", 0))


def controllers_to_moves(controllers):
    """Pack a controller list into one [5] row of `chromo_move_state`; move types
    that have no controller are off."""
    row = np.zeros(NUM_MOVES, dtype=MOVE_DTYPE)
    row["alpha"] = 2 / 21
    seen = set()
    for c in controllers:
        m = c.move
        if m.name not in MOVE_ID:
            raise NotImplementedError(f"move {m.name} is outside the accelerated path")
        i = MOVE_ID[m.name]
        if i in seen:
            raise NotImplementedError("one controller per move type")
        seen.add(i)
        r = row[i]
        r["amp_move"], r["amp_bead"] = m.amp_move, m.amp_bead
        r["move_amp_lo"], r["move_amp_hi"] = c.move_amp_bounds
        r["bead_amp_lo"], r["bead_amp_hi"] = c.bead_amp_bounds
        r["acceptance_rate"] = m.acceptance_tracker.acceptance_rate
        r["alpha"] = m.acceptance_tracker.alpha
        r["num_attempt"], r["num_success"] = m.num_attempt, m.num_success
        r["num_per_cycle"], r["move_on"] = m.num_per_cycle, int(m.move_on)
        r["controller"] = getattr(c, "device_code", 0)
    return row


def controller_order(controllers):
    """Move ids in the order the reference would go through them within one MC step: its controller list as
    given (`for controller in mc_move_controllers`, mc_sim.pyx:92-103), then the move types without a
    controller (they are off)."""
    order = [MOVE_ID[c.move.name] for c in controllers]
    return order + [i for i in range(NUM_MOVES) if i not in order]


def moves_to_controllers(row, controllers):
    for c in controllers:
        r = row[MOVE_ID[c.move.name]]
        m = c.move
        m.amp_move, m.amp_bead = float(r["amp_move"]), int(r["amp_bead"])
        m.num_attempt, m.num_success = int(r["num_attempt"]), int(r["num_success"])
        m.acceptance_tracker.acceptance_rate = float(r["acceptance_rate"])


def mc_sim(polymers, readerproteins, num_mc_steps, mc_move_controllers, field, mu_adjust_factor,
           random_seed):
    """Perform `num_mc_steps` Monte-Carlo sweeps (mc_sim.pyx:26-103)."""
    if len(polymers) != 1:
        raise NotImplementedError("one polymer per field (fields.pyx:53-59); use ReplicaEnsemble for replicas")
    poly = polymers[0]
    if poly not in field:
        # mc_sim.pyx:144: a polymer outside the field feels no field at all
        from ..fields import NullField
        field = poly.__dict__.setdefault("_null_field", NullField([]))
    poly.mu_adjust_factor = mu_adjust_factor
    poly._field = field
    e = field._push(poly)
    mv = controllers_to_moves(mc_move_controllers).reshape(1, NUM_MOVES).copy()
    e.set_move_order(controller_order(mc_move_controllers))
    mode = rng_mode()
    e.mc_sim(int(num_mc_steps), mv, float(mu_adjust_factor), int(random_seed), mode,
             numpy_seeds=int(random_seed) & 0xFFFFFFFF if mode == RNG_REPLAY else None)
    field._pull(poly, e)
    moves_to_controllers(mv[0], mc_move_controllers)


def mc_step(adaptible_move, poly, readerproteins, field, active_field, update_distances=False):
    """One move attempt (mc_sim.pyx:106-182) through the same device code path."""
    if not (poly in field and active_field):
        from ..fields import NullField
        field = poly.__dict__.setdefault("_null_field", NullField([]))
    poly._field = field
    e = field._push(poly)
    out = e.mc_step(0, MOVE_ID[adaptible_move.name], adaptible_move.amp_move, adaptible_move.amp_bead,
                    float(poly.mu_adjust_factor), rng_mode(), _seed_for(poly), force_accept=-1)
    adaptible_move.num_attempt += 1
    if len(out["inds"]) == 0:
        return
    field._pull(poly, e)
    if out["accepted"]:
        adaptible_move.num_success += 1
    adaptible_move.acceptance_tracker.update_acceptance_rate(1.0 if out["accepted"] else 0.0, False)
    poly._last_step = out
