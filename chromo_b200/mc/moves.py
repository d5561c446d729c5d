"""Move adapter (host mirror of chromo/mc/moves.pyx: MCAdapter 58-299, Bounds
302-342).  On the device the same state is `chromo_move_state`."""
import numpy as np
import pandas as pd

from ..util import mc_stat
from .move_funcs import change_binding_state, crank_shaft, end_pivot, slide, tangent_rotation


class MCAdapter:
    """Tracks attempts / successes / amplitudes of one move type."""

    def __init__(self, log_dir, log_file_prefix, move_func, moves_in_average, init_amp_bead, init_amp_move):
        self.name = move_func.__name__
        self.move_func = move_func
        self.amp_move = float(init_amp_move)
        self.num_per_cycle = 1
        if int(init_amp_bead) != init_amp_bead:
            raise TypeError("'float' object cannot be interpreted as an integer")  # `long` argument
        self.amp_bead = int(init_amp_bead)
        self.num_attempt = 0
        self.num_success = 0
        self.move_on = 1
        self.last_amp_move = 0
        self.last_amp_bead = 0
        self.acceptance_tracker = mc_stat.AcceptanceTracker(log_dir, log_file_prefix, float(moves_in_average))

    def __str__(self):
        return f"MCAdapter<{self.name}>"

    def to_file(self, path):
        pass

    def propose(self, polymer):
        """moves.pyx:137-154."""
        self.num_attempt += 1
        return self.move_func(polymer=polymer, amp_move=self.amp_move, amp_bead=self.amp_bead)

    def accept(self, poly, dE, inds, n_inds, log_move, log_update, update_distances):
        """Copy the trial state over the current one (moves.pyx:156-239)."""
        inds = np.asarray(inds[:n_inds])
        if self.name == "change_binding_state":
            poly.states[inds] = poly.states_trial[inds]
        elif self.name == "slide":
            poly.r[inds] = poly.r_trial[inds]
            poly.t3_trial[inds] = poly.t3[inds]
            poly.t2_trial[inds] = poly.t2[inds]
        elif self.name == "tangent_rotation":
            poly.t3[inds] = poly.t3_trial[inds]
            poly.t2[inds] = poly.t2_trial[inds]
            poly.r_trial[inds] = poly.r[inds]
        else:
            poly.r[inds] = poly.r_trial[inds]
            poly.t3[inds] = poly.t3_trial[inds]
            poly.t2[inds] = poly.t2_trial[inds]
        self.num_success += 1
        self.acceptance_tracker.update_acceptance_rate(accept=1.0, log_update=log_update)
        if log_move == 1:
            self.acceptance_tracker.log_move(self.amp_move, self.amp_bead, poly.last_amp_move,
                                             poly.last_amp_bead, dE)

    def reject(self, poly, dE, inds, n_inds, log_move, log_update, update_distances):
        """Reset the trial state (moves.pyx:241-299)."""
        inds = np.asarray(inds[:n_inds])
        if self.name == "change_binding_state":
            poly.states_trial[inds] = poly.states[inds]
        else:
            poly.r_trial[inds] = poly.r[inds]
            poly.t3_trial[inds] = poly.t3[inds]
            poly.t2_trial[inds] = poly.t2[inds]
        self.acceptance_tracker.update_acceptance_rate(accept=0.0, log_update=log_update)
        if log_move == 1:
            self.acceptance_tracker.log_move(self.amp_move, self.amp_bead, poly.last_amp_move,
                                             poly.last_amp_bead, dE)


class Bounds:
    """Named move / bead amplitude bounds (moves.pyx:302-342)."""

    def __init__(self, name, bounds):
        self.name = name
        self.bounds = bounds

    def to_dataframe(self):
        move_names = self.bounds.keys()
        arr = np.atleast_2d(np.array(list(self.bounds.values())).flatten())
        cols = pd.MultiIndex.from_product([move_names, ('lower_bound', 'upper_bound')])
        return pd.DataFrame(arr, columns=cols)

    def to_csv(self, path):
        return self.to_dataframe().to_csv(path)

    def to_file(self, path):
        return self.to_csv(path)


move_list = [crank_shaft, end_pivot, slide, tangent_rotation, change_binding_state]
