#!/usr/bin/env python
"""Full density recompute (A8, update_all_densities fields.pyx:1977-2106) at C2 / C3 / C1 sizes: CUDA-event time
of chromo_field_recompute, algorithmic bytes (24 N + nb N + 8 (nb+1) n_bins per replica) / time against the
measured copy bandwidth, and a bit-reproducibility check (two recomputes, identical bytes)."""
import json, math, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import bench
from chromo_b200.ensemble import ReplicaEnsemble

peak = 6650.0
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass
out = []
for label, R, N, nb in (("C2", 1024, 10000, 1), ("C3", 1024, 10000, 2), ("C1-sized", 1024, 1000, 1)):
    r, t3, t2, states, mods, grid = bench.make_inputs(R, N, 5, pinned=False, nb=nb)
    if N == 1000:
        grid = dict(grid, nx=16, ny=16, nz=16)
    binders = [dict(bench.HP1)] + ([dict(bench.HP1, name="PRC1", cross_talk={})] if nb == 2 else [])
    states = np.random.default_rng(1).integers(0, 3, size=states.shape)
    ens = ReplicaEnsemble(r, t3, t2, states, mods, binders=binders, bond_params=bench.bond_params(N), grid=grid,
                          bead_vol=(4 / 3) * math.pi * 125.0, moves=bench.stationary_moves(R, N))
    eng = ens.engine
    stream = torch.cuda.ExternalStream(eng.stream())
    for _ in range(3):
        eng.field_recompute(clamp=True)
    eng.sync()
    d1 = ens.density().copy()
    K = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(K):
        eng.field_recompute(clamp=True)
    e1.record(stream)
    eng.sync()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    d2 = ens.density()
    n_bins = grid["nx"] ** 3
    algo = R * (24 * N + nb * N + 8 * (nb + 1) * n_bins)
    out.append(dict(case=label, replicas=R, beads=N, binders=nb, voxels=n_bins, ms=ms, algorithmic_bytes=algo,
                    achieved_gbs=algo / (ms * 1e-3) / 1e9, frac_of_measured_copy_peak=algo / (ms * 1e-3) / 1e9 / peak,
                    beads_per_s=R * N / (ms * 1e-3), bit_reproducible=bool(d1.tobytes() == d2.tobytes()),
                    mass_ok=bool(np.allclose(d2[..., 0].sum(axis=1) * (grid["x_width"] ** 3 / n_bins), N, rtol=1e-12))))
    print(json.dumps(out[-1]), flush=True)
    ens.close()
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(json.dumps(out, indent=1))
