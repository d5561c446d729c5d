#!/usr/bin/env python
"""Phi = sum_bins (V/v) phi^2 of the C2 chromatin at several chi (mean / spread over replicas) and the swap
acceptance of candidate ladders: input for the spacing of the C5 ladder (20-40 % acceptance wanted)."""
import json, math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import bench
from chromo_b200 import parallel as par
from chromo_b200.ensemble import ReplicaEnsemble

R, N = 512, 10000
r, t3, t2, states, mods, grid = bench.make_inputs(R, N, 77, pinned=False)
for L, lo, hi in ((32, 0.5, 2.0), (16, 0.5, 2.0), (16, 0.25, 4.0), (8, 0.5, 2.0)):
    ladder = np.tile(np.geomspace(lo, hi, L), R // L)
    ens = ReplicaEnsemble(r.copy(), t3.copy(), t2.copy(), states.copy(), mods, binders=[dict(bench.HP1)],
                          bond_params=bench.bond_params(N), grid=grid, bead_vol=(4 / 3) * math.pi * 125.0, chi=ladder,
                          mu=[-1.2], moves=bench.stationary_moves(R, N))
    ex = par.ReplicaExchange(ens, ladder, seed=5, ladder_len=L)
    ens.mc_sim(200, 1.0, 1, sync_host=False)
    _, _, t0, a0 = ex.state()
    phis = []
    for rnd in range(20):
        ens.mc_sim(10, 1.0, 10 + rnd, sync_host=False)
        ex.step()
    rung, chi, tried, acc = ex.state()
    phi = ex.phi_all.cpu().numpy()
    byr = np.stack([phi[rung[l0:l0 + L]] for l0 in range(0, R, L)])
    print(json.dumps(dict(L=L, lo=lo, hi=hi, acceptance=acc / max(1, tried), phi_mean_lo=float(byr[:, 0].mean()),
                          phi_mean_hi=float(byr[:, -1].mean()), phi_std_same_rung=float(byr.std(axis=0).mean()))), flush=True)
    ens.close()
