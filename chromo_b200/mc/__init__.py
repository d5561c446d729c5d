"""Monte-Carlo driver (host mirror of chromo/mc/__init__.py).

`polymer_in_field` keeps the reference's snapshot loop (mc/__init__.py:36-160):
per snapshot an annealing factor, an inner seed, one `mc_sim` call, then the
configuration is handed to `save_snapshot`.  Output-directory bookkeeping
(`make_reproducible`) is out of scope.
"""
from pathlib import Path
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .mc_controller import Controller, SimpleControl, all_moves
from .mc_sim import mc_sim, mc_step, set_rng_mode, rng_mode  # noqa: F401
from .moves import Bounds


def get_amplitude_bounds(polymers) -> Tuple[Bounds, Bounds]:
    """Lower / upper bounds of the bead-selection and move amplitudes
    (mc/__init__.py:295-332)."""
    poly_len = np.min([polymer.r.shape[0] for polymer in polymers])
    min_spacing = np.min([np.min(polymer.bead_length) for polymer in polymers])
    bead_amp_bounds = Bounds("bead_amp_bounds", {
        "crank_shaft": (min(30, poly_len), min(150, poly_len)),
        "slide": (min(10, poly_len), min(150, poly_len)),
        "end_pivot": (min(50, poly_len / 4), min(150, int(poly_len / 2))),
        "tangent_rotation": (1, poly_len),
        "change_binding_state": (1, 1),
    })
    move_amp_bounds = Bounds("move_amp_bounds", {
        "crank_shaft": (0.1 * np.pi, 0.25 * np.pi),
        "slide": (0.2 * min_spacing, 0.3 * min_spacing),
        "end_pivot": (0.2 * np.pi, 0.25 * np.pi),
        "tangent_rotation": (0.05 * np.pi, 0.2 * np.pi),
        "change_binding_state": (0, 0),
    })
    return bead_amp_bounds, move_amp_bounds


def polymer_in_field(polymers, binders, field, num_save_mc, num_saves, bead_amp_bounds, move_amp_bounds,
                     mc_move_controllers: Optional[List[Controller]] = None, random_seed: Optional[int] = 0,
                     mu_schedule=None, output_dir: Optional[str] = '.',
                     save_snapshot: Optional[Callable] = None, **kwargs):
    """Monte-Carlo simulation of polymers in a field, `num_saves` snapshots of
    `num_save_mc` sweeps each (mc/__init__.py:36-160)."""
    np.random.seed(random_seed)
    if mc_move_controllers is None:
        mc_move_controllers = all_moves(log_dir=output_dir, bead_amp_bounds=bead_amp_bounds.bounds,
                                        move_amp_bounds=move_amp_bounds.bounds, controller=SimpleControl)
    for mc_count in range(num_saves):
        if mu_schedule is not None:
            mu_adjust_factor = mu_schedule.function(mc_count, num_saves)
        else:
            mu_adjust_factor = 1
        inner_seed = np.random.randint(0, 1E9)
        mc_sim(polymers, binders, num_save_mc, mc_move_controllers, field, mu_adjust_factor, inner_seed)
        if save_snapshot is not None:
            save_snapshot(mc_count, polymers, field, mc_move_controllers)
    return mc_move_controllers


_polymer_in_field = polymer_in_field
