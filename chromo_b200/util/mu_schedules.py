"""Simulated-annealing schedules: the factor applied to the chemical potential
at snapshot i of N (a host scalar -> kernel argument `mu_adjust_factor`).
Same names and call convention as chromo/util/mu_schedules.py (Schedule 17-31,
the schedules 34-164), expressed through one piecewise-linear helper."""
import numpy as np


class Schedule:
    """Wraps f(i, N) -> float; `mc.polymer_in_field` calls `.function`."""

    def __init__(self, fxn):
        self.name = fxn.__name__
        self.function = fxn

    def to_file(self, path):
        with open(path, "w") as f:
            f.write(self.name)


def _ramp(i, N, hi, lo, start, end):
    """hi while i <= start, lo once i >= N - end, linear in between."""
    if i <= start:
        return float(hi)
    if end is not None and i >= N - end:
        return float(lo)
    span = float(N - (start + (end or 0)))
    return (lo - hi) * (float(i - start) / span) + hi


def linear_1(i, N):
    return 2 * (i / N) - 1


def linear_2_for_negative_cp(i, N):
    """4 -> 1, held at 4 for the first quarter (mu_schedules.py:131-138)."""
    return _ramp(i, N, 4., 1., np.floor(N / 4), None)


def linear_step_for_negative_cp(i, N):
    """4 -> 1 with a fifth of the snapshots held at each end (mu_schedules.py:141-151)."""
    return _ramp(i, N, 4., 1., np.floor(N / 5), np.floor(N / 5))


def linear_step_for_negative_cp_mild(i, N):
    """2 -> 1 with a fifth of the snapshots held at each end (mu_schedules.py:154-164)."""
    return _ramp(i, N, 2., 1., np.floor(N / 5), np.floor(N / 5))


def constant(i, N):
    return 1.0
