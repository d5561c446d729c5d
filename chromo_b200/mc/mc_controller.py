"""Amplitude controllers (chromo/mc/mc_controller.py: Controller 16-69,
NoControl 72-88, SimpleControl 91-213, controller-list builders 216-468).

During `mc_sim` the controller runs IN the kernel once per MC step and move type
(McWarp::update_amplitudes); these host classes hold the same state between
calls and apply the same rule when driven from Python."""
from abc import ABC, abstractmethod

from . import moves as mv


class Controller(ABC):
    def __init__(self, mc_adapter, bead_amp_bounds, move_amp_bounds):
        if bead_amp_bounds[0] > bead_amp_bounds[1]:
            raise ValueError("Lower bead amplitude bound must be less than upper bead amplitude bound")
        if move_amp_bounds[0] > move_amp_bounds[1]:
            raise ValueError("Lower move amplitude bound must be less than upper move amplitude bound")
        self.move = mc_adapter
        self.bead_amp_bounds = bead_amp_bounds
        self.move_amp_bounds = move_amp_bounds

    device_code = 0

    @abstractmethod
    def update_amplitudes(self):
        pass


class NoControl(Controller):
    device_code = 0

    def update_amplitudes(self):
        return

    def update_move_amplitude(self):
        return

    def update_bead_amplitude(self):
        return


class SimpleControl(Controller):
    """Scale the move amplitude by 0.95 / (1/0.95) towards a 0.5 acceptance rate;
    at a bound, step the bead amplitude by one and reset the move amplitude to the
    opposite bound (mc_controller.py:148-213)."""
    device_code = 1

    def __init__(self, mc_adapter, bead_amp_bounds, move_amp_bounds):
        super().__init__(mc_adapter, bead_amp_bounds, move_amp_bounds)
        self.name = "SimpleController"

    def to_file(self, path):
        pass

    def update_amplitudes(self, setpoint_acceptance=0.5, move_adjust_factor=0.95, num_delta_beads=1):
        self.update_move_amplitude(setpoint_acceptance, move_adjust_factor, num_delta_beads)

    def update_move_amplitude(self, setpoint_acceptance=0.5, move_adjust_factor=0.95, num_delta_beads=1):
        acceptance = self.move.acceptance_tracker.acceptance_rate
        if acceptance < setpoint_acceptance:
            prop = self.move.amp_move * move_adjust_factor
            if prop > self.move_amp_bounds[0]:
                self.move.amp_move = prop
            else:
                self.move.amp_bead = max(self.bead_amp_bounds[0],
                                         self.update_bead_amplitude(False, num_delta_beads))
        elif acceptance > setpoint_acceptance:
            prop = self.move.amp_move / move_adjust_factor
            if prop < self.move_amp_bounds[1]:
                self.move.amp_move = prop
            else:
                self.move.amp_bead = min(self.bead_amp_bounds[1],
                                         self.update_bead_amplitude(True, num_delta_beads))

    def update_bead_amplitude(self, increase, num_delta_beads=1):
        if increase:
            self.move.amp_move = self.move_amp_bounds[0]
            return self.move.amp_bead + num_delta_beads
        self.move.amp_move = self.move_amp_bounds[1]
        return self.move.amp_bead - num_delta_beads


def _controllers(move_fxns, log_dir, bead_amp_bounds, move_amp_bounds, controller, per_cycle):
    out = [
        controller(
            mv.MCAdapter(str(log_dir) + '/acceptance_trackers', move.__name__ + "_snap_", move,
                         moves_in_average=20, init_amp_bead=bead_amp_bounds[move.__name__][0],
                         init_amp_move=move_amp_bounds[move.__name__][0]),
            bead_amp_bounds=bead_amp_bounds[move.__name__],
            move_amp_bounds=move_amp_bounds[move.__name__])
        for move in move_fxns]
    for c, k in zip(out, per_cycle):
        c.move.num_per_cycle = k
    return out


def all_moves(log_dir, bead_amp_bounds, move_amp_bounds, controller=NoControl):
    """The canonical 161-attempt sweep: 30 crank-shaft, 1 end-pivot, 60 slide,
    60 tangent-rotation, 10 binding (mc_controller.py:216-266)."""
    return _controllers(mv.move_list, log_dir, bead_amp_bounds, move_amp_bounds, controller,
                        (30, 1, 60, 60, 10))


def all_moves_except_binding_state(log_dir, bead_amp_bounds, move_amp_bounds, controller=NoControl):
    """mc_controller.py:269-320."""
    return _controllers(mv.move_list[:4], log_dir, bead_amp_bounds, move_amp_bounds, controller, (5, 5, 5, 5))


def specific_move(move_fxn, log_dir, bead_amp_bounds, move_amp_bounds, controller=NoControl):
    """mc_controller.py:323-369."""
    return _controllers([move_fxn], log_dir, bead_amp_bounds, move_amp_bounds, controller, (1,))


def specific_moves(move_fxns, log_dir, bead_amp_bounds, move_amp_bounds, controller=NoControl):
    """mc_controller.py:372-420."""
    return _controllers(list(move_fxns), log_dir, bead_amp_bounds, move_amp_bounds, controller,
                        (1,) * len(move_fxns))


def only_binding_move(log_dir, bead_amp_bounds, move_amp_bounds, controller=NoControl):
    """mc_controller.py:423-468."""
    return _controllers([mv.change_binding_state], log_dir, bead_amp_bounds, move_amp_bounds, controller, (1,))
