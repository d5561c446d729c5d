#!/usr/bin/env python
"""Cycles per phase of an MC attempt, per move type (development tool).

    tools/build_variant.sh WORK /tmp/timers.so -DCB_PHASE_TIMERS
    python tools/phase_timers.py --lib /tmp/timers.so [--warps 1,2]

The instrumented build accumulates clock64() deltas per warp (mc_kernel.cuh, CB_PHASE_TIMERS);
this script runs the bench workload (C2) at the controllers' working point and prints, per
move type, the mean cycles per attempt spent in each phase, per warp that executed it.
"""
import argparse
import ctypes as C
import json
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import bench  # noqa: E402
from chromo_b200 import _lib  # noqa: E402
from chromo_b200._lib import MOVE_NAMES  # noqa: E402
from chromo_b200.ensemble import ReplicaEnsemble, default_moves  # noqa: E402

PH = ["prefetch", "stage1", "wait_turn", "stale/redo", "energy", "metropolis", "commit", "pass+clear", "prepare",
      "blk_sync", "type_sync", "attempts", "s1:final", "s1:dEpoly", "s1:scatter", "n_sum"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", required=True)
    ap.add_argument("--replicas", type=int, default=1024)
    ap.add_argument("--beads", type=int, default=10000)
    ap.add_argument("--warm", type=int, default=100)
    ap.add_argument("--sweeps", type=int, default=20)
    ap.add_argument("--warps", default="1,2")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    sys.path.insert(0, str(ROOT / "tests"))
    import devlib
    devlib.use_library(a.lib)
    L = C.CDLL(a.lib)
    R, N = a.replicas, a.beads
    r, t3, t2, states, mods, grid = bench.make_inputs(R, N, 1234, pinned=False)
    res = {}
    for w in [int(x) for x in a.warps.split(",")]:
        ens = ReplicaEnsemble(r.copy(), t3.copy(), t2.copy(), states.copy(), mods, binders=[dict(bench.HP1)],
                              bond_params=bench.bond_params(N), grid=grid, bead_vol=(4 / 3) * math.pi * 5.0 ** 3,
                              chi=1.0, mu=[-1.2], moves=bench.stationary_moves(R, N), device=0)
        eng = ens.engine
        eng.set_warps_per_replica(w)
        cap = eng.set_table_capacity(0)
        ens.mc_sim(a.warm, 1.0, 99, sync_host=False)
        ens.sync()
        buf = (C.c_ulonglong * (5 * len(PH)))()
        assert L.cb_phase_read(buf) == 0
        ens.mc_sim(a.sweeps, 1.0, 7, sync_host=False)
        ens.sync()
        assert L.cb_phase_read(buf) == 0
        t = np.array(list(buf), dtype=np.float64).reshape(5, len(PH))
        print(f"\nwarps per replica = {w}, table slots = {cap}: cycles per attempt (mean over warps that ran it)")
        print("%-22s" % "" + " ".join("%10s" % p[:10] for p in PH[:11]) + "      total")
        out = {}
        for m, name in enumerate(MOVE_NAMES):
            n = max(t[m, 11], 1.0)
            row = t[m, :11] / n
            sub = t[m, 12:] / n
            out[name] = dict(zip(PH[:11] + PH[12:], [round(float(x), 1) for x in list(row) + list(sub)]),
                             attempts=int(t[m, 11]))
            print("%-22s" % name + " ".join("%10.0f" % x for x in row) + "%11.0f" % row.sum() +
                  "   | " + " ".join("%s=%.0f" % (k, v) for k, v in zip(PH[12:], sub)))
        res[f"w{w}"] = out
        ens.close()
    if a.out:
        Path(a.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
