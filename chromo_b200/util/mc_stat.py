"""Acceptance tracking (host mirror of chromo/util/mc_stat.py:51-207).  On the
device the same EWMA lives in `chromo_move_state.acceptance_rate`."""
import csv
import os
from typing import List


class AcceptanceTracker:
    def __init__(self, log_dir: str, log_file_prefix: str, moves_in_average: float):
        self.log_dir = log_dir
        self.log_file_prefix = log_file_prefix
        self.alpha = 2 / (moves_in_average + 1)  # mc_stat.py:72
        self.acceptance_rate = 0
        self.amp_bead_limit_log: List[float] = []
        self.amp_move_limit_log: List[float] = []
        self.amp_bead_realized_log: List[float] = []
        self.amp_move_realized_log: List[float] = []
        self.dE_log = []
        self.move_accepted = []
        self.acceptance_log: List[float] = []

    def create_log_file(self, ind: int):
        os.makedirs(self.log_dir, exist_ok=True)
        log_path = self.log_dir + "/" + self.log_file_prefix + str(ind) + ".csv"
        with open(log_path, "w") as f:
            csv.writer(f).writerow(["snapshot", "iteration", "bead_amp_limit", "move_amp_limit",
                                    "bead_amp_realized", "move_amp_realized", "dE", "accepted",
                                    "acceptance_rate"])

    def log_move(self, amp_move_limit, amp_bead_limit, amp_move, amp_bead, dE):
        self.amp_move_limit_log.append(amp_move_limit)
        self.amp_bead_limit_log.append(amp_bead_limit)
        self.amp_move_realized_log.append(amp_move)
        self.amp_bead_realized_log.append(amp_bead)
        self.dE_log.append(dE)
        self.acceptance_log.append(self.acceptance_rate)

    def save_move_log(self, snapshot: int):
        log_path = self.log_dir + "/" + self.log_file_prefix + str(snapshot) + ".csv"
        with open(log_path, 'a') as output:
            w = csv.writer(output, delimiter=',')
            for i in range(len(self.amp_move_realized_log)):
                w.writerow([snapshot, i + 1, self.amp_bead_limit_log[i], self.amp_move_limit_log[i],
                            self.amp_bead_realized_log[i], self.amp_move_realized_log[i], self.dE_log[i],
                            self.move_accepted[i], self.acceptance_log[i]])
        self.amp_move_limit_log, self.amp_bead_limit_log = [], []
        self.amp_move_realized_log, self.amp_bead_realized_log = [], []
        self.dE_log, self.move_accepted, self.acceptance_log = [], [], []

    def update_acceptance_rate(self, accept: float, log_update: int):
        """EWMA update, mc_stat.py:190-207."""
        if log_update == 1:
            self.move_accepted.append(accept)
        self.acceptance_rate = (self.alpha * accept) + (1 - self.alpha) * self.acceptance_rate
