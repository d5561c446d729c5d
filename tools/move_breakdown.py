#!/usr/bin/env python
"""Per-move-type cost of the MC kernel at the controllers' working point.

    python tools/move_breakdown.py [--replicas 1024] [--beads 10000] [--warm 400] [--sweeps 20]

Runs the bench workload (C2) until SimpleControl has settled, freezes the
amplitudes, then times `mc_sim` with ONE move type switched on at a time (CUDA
events on the context's stream).  Prints microseconds per attempt per replica
and each type's share of a canonical 161-attempt sweep -- the number that says
which part of the kernel to work on next.  Development tool, not a bench line.
"""
import argparse
import json
import math
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from chromo_b200._lib import MOVE_NAMES  # noqa: E402
from chromo_b200.ensemble import ReplicaEnsemble, default_moves  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=1024)
    ap.add_argument("--beads", type=int, default=10000)
    ap.add_argument("--warm", type=int, default=100)
    ap.add_argument("--sweeps", type=int, default=20)
    ap.add_argument("--table-slots", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0, help="warps per replica (0 = library default)")
    ap.add_argument("--rpb", type=int, default=0, help="replicas per block (0 = library default)")
    ap.add_argument("--lib", default=None, help="another build of the C-ABI library (tools/build_variant.sh), for A/B runs")
    a = ap.parse_args()
    R, N = a.replicas, a.beads
    if a.lib:
        from chromo_b200 import _lib
        sys.path.insert(0, str(ROOT / "tests"))
        import devlib
        devlib.use_library(a.lib)
    r, t3, t2, states, mods, grid = bench.make_inputs(R, N, 1234, pinned=False)
    ens = ReplicaEnsemble(r, t3, t2, states, mods, binders=[dict(bench.HP1)], bond_params=bench.bond_params(N),
                          grid=grid, bead_vol=(4 / 3) * math.pi * 5.0 ** 3, chi=1.0, mu=[-1.2],
                          moves=bench.stationary_moves(R, N), device=0)
    eng = ens.engine
    warps = eng.set_warps_per_replica(a.warps)
    rpb = eng.set_replicas_per_block(a.rpb) if hasattr(eng._L, "chromo_ctx_set_replicas_per_block") else 1
    cap = eng.set_table_capacity(a.table_slots)
    stream = torch.cuda.ExternalStream(eng.stream(), device=0)

    def timed(sweeps, seed):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.sync()
        e0.record(stream)
        ens.mc_sim(sweeps, 1.0, seed, sync_host=False)
        e1.record(stream)
        eng.sync()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    ens.mc_sim(a.warm, 1.0, 99, sync_host=False)
    ens.sync()
    base = ens.moves.copy()
    per_cycle = base["num_per_cycle"][0].copy()
    out = dict(warps=warps, rpb=rpb, table_slots=cap, amp_bead_mean=[round(float(x), 2) for x in base["amp_bead"].mean(axis=0)],
               amp_move_mean=[round(float(x), 4) for x in base["amp_move"].mean(axis=0)])
    timed(2, 5)
    ms_all = timed(a.sweeps, 7)
    out["all"] = dict(ms=ms_all, us_per_attempt=1e3 * ms_all / (a.sweeps * per_cycle.sum()),
                      attempts_per_s=R * a.sweeps * int(per_cycle.sum()) / (ms_all * 1e-3))
    tot = 0.0
    for i, name in enumerate(MOVE_NAMES):
        m = base.copy()
        m["controller"][:] = 0
        m["move_on"][:] = 0
        m["move_on"][:, i] = 1
        eng.set_moves(m)
        timed(2, 11 + i)
        ms = timed(a.sweeps, 21 + i)
        us = 1e3 * ms / (a.sweeps * int(per_cycle[i]))
        out[name] = dict(ms=ms, us_per_attempt=us, sweep_us=us * int(per_cycle[i]))
        tot += us * int(per_cycle[i])
    for name in MOVE_NAMES:
        out[name]["share"] = out[name]["sweep_us"] / tot
    out["sum_sweep_us"] = tot
    print(json.dumps(out, indent=1))
    ens.close()


if __name__ == "__main__":
    main()
