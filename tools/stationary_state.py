#!/usr/bin/env python
"""Where SimpleControl takes the bench workload (C2) and how long it needs: amplitudes per move type
(mean / spread over replicas) every `--chunk` sweeps.  The bench starts both arms from the state this
finds (profiles/stationary_amplitudes.json) so that they do equal work from the first timed step.

    python tools/stationary_state.py --sweeps 8000 --chunk 500 --out profiles/stationary_amplitudes.json
"""
import argparse
import json
import math
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from chromo_b200._lib import MOVE_NAMES  # noqa: E402
from chromo_b200.ensemble import ReplicaEnsemble, default_moves  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=1024)
    ap.add_argument("--beads", type=int, default=10000)
    ap.add_argument("--sweeps", type=int, default=8000)
    ap.add_argument("--chunk", type=int, default=500)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    R, N = a.replicas, a.beads
    r, t3, t2, states, mods, grid = bench.make_inputs(R, N, 1234, pinned=False)
    ens = ReplicaEnsemble(r, t3, t2, states, mods, binders=[dict(bench.HP1)], bond_params=bench.bond_params(N),
                          grid=grid, bead_vol=(4 / 3) * math.pi * 5.0 ** 3, chi=1.0, mu=[-1.2],
                          moves=default_moves(R, N, 16.5), device=0)
    hist = []
    done = 0
    while done < a.sweeps:
        t0 = time.perf_counter()
        ens.mc_sim(a.chunk, 1.0, 99 + done, sync_host=False)
        ens.sync()
        dt = time.perf_counter() - t0
        done += a.chunk
        mv = ens.moves
        row = dict(sweeps=done, attempts_per_s=R * a.chunk * 161 / dt,
                   amp_bead_mean=[float(x) for x in mv["amp_bead"].mean(axis=0)],
                   amp_bead_std=[float(x) for x in mv["amp_bead"].std(axis=0)],
                   amp_move_mean=[float(x) for x in mv["amp_move"].mean(axis=0)],
                   acceptance=[float(x) for x in mv["acceptance_rate"].mean(axis=0)],
                   bound_frac=float(ens.states_mean()) if hasattr(ens, "states_mean") else None)
        hist.append(row)
        print(json.dumps(row), flush=True)
    mv = ens.moves
    out = dict(workload=f"C2: {R} replicas x {N} beads, HP1, SimpleControl from get_amplitude_bounds' lower bounds",
               sweeps=done, move_names=list(MOVE_NAMES),
               amp_bead=[int(round(float(x))) for x in mv["amp_bead"].mean(axis=0)],
               amp_move=[float(x) for x in mv["amp_move"].mean(axis=0)],
               acceptance_rate=[float(x) for x in mv["acceptance_rate"].mean(axis=0)],
               history=hist)
    if a.out:
        Path(a.out).write_text(json.dumps(out, indent=1))
    ens.close()


if __name__ == "__main__":
    main()
