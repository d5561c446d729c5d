#!/usr/bin/env python
"""Time the host-array path (chromo_mc_sim_host) for several chunk counts (development tool)."""
import math, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np, torch
import bench
from chromo_b200.ensemble import ReplicaEnsemble, default_moves
R, N, S = 1024, 10000, int(sys.argv[1]) if len(sys.argv) > 1 else 50
r, t3, t2, states, mods, grid = bench.make_inputs(R, N, 1234, pinned=True)
ens = ReplicaEnsemble(r, t3, t2, states, mods, binders=[dict(bench.HP1)], bond_params=bench.bond_params(N), grid=grid,
                      bead_vol=(4 / 3) * math.pi * 5.0 ** 3, chi=1.0, mu=[-1.2], moves=default_moves(R, N, 16.5), device=0)
ens.mc_sim(400, 1.0, 99, sync_host=False); ens.sync()
ens.moves["controller"][:] = 0  # freeze the amplitudes: every call below does the same work
ens.engine.set_moves(ens.moves)
for ch in (-1, 1, 2, 4, 3, -1, 4, 2):
    ens.mc_sim(S, 1.0, 5, n_chunks=ch)
    t0 = time.perf_counter()
    for k in range(3):
        ens.mc_sim(S, 1.0, 6 + k, n_chunks=ch)
    dt = (time.perf_counter() - t0) / 3
    print(f"chunks={ch:2d}  {dt*1e3:7.1f} ms per call  {R*S*161/dt/1e6:7.1f} M attempts/s", flush=True)
t0 = time.perf_counter(); ens.push(); t1 = time.perf_counter(); ens.pull(); t2_ = time.perf_counter()
print(f"push {1e3*(t1-t0):.1f} ms, pull {1e3*(t2_-t1):.1f} ms")
ens.close()
