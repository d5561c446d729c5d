#!/usr/bin/env python
"""Bench lines for the BASELINE configurations other than the headline one (bench.py measures C2):

  c1      homopolymer SSWLC, 1,000 beads, `null_reader` (no binders), PERIODIC box (confine_type = ""),
          16^3 voxels of 28.6 nm, 1,024 replicas                                  (BASELINE configs[0])
  c3      chromatin 10,000 beads, HP1 + PRC1 on the H3K9me3 / H3K27me3 tracks with cross-talk, chemical
          potentials swept over linspace(-2, 0) across 1,024 replicas              (configs[2])
  c4      chromosome-scale chromatin, 400,000 beads, 65^3 voxels, 148 replicas (one per SM)   (configs[3])
  c4fine  the same on the fine 130^3 grid (2.2 M voxels per replica: where the reference's two
          O(n_bins) scans per move, fields.pyx:1223-1226, 1971-1975, cost milliseconds per attempt)

    python tools/bench_configs.py c1 c3 c4 c4fine --out-dir profiles

Each line: attempts/s with the state resident (CUDA events around K back-to-back mc_sim launches after W
warm-up launches under SimpleControl from the amplitude bounds' lower ends), algorithmic bytes / time against
the measured copy bandwidth, acceptance, bead windows.  Parity for these shapes: tests/test_gpu_scale.py."""
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
from chromo_b200.ensemble import ReplicaEnsemble, default_moves  # noqa: E402
from chromo_b200.util import chemical_mods  # noqa: E402
from chromo_b200.util import poly_paths as paths  # noqa: E402

NULL = dict(name="null_reader", sites_per_bead=0, bind_energy_mod=0.0, bind_energy_no_mod=0.0, interaction_energy=0.0,
            chemical_potential=0.0, interaction_radius=0.0, cross_talk={})
PRC1 = dict(name="PRC1", sites_per_bead=2, bind_energy_mod=-0.01, bind_energy_no_mod=1.52, interaction_energy=-4.0,
            chemical_potential=-1.2, interaction_radius=3.0, cross_talk={"HP1": 0.0})


def build(name, seed=0):
    rng = np.random.default_rng(seed)
    if name == "c1":
        R, N, sweeps = 1024, 1000, 200
        W, nx = 16 * 28.6, 16
        r = paths.gaussian_walk(N - 1, np.full(N - 1, 16.5), rng, replicas=R)
        binders, tracks, mu, confine, Rc = [dict(NULL)], None, [0.0], "", 0.0
        label = "C1: homopolymer SSWLC 1,000 beads, null_reader, periodic box 16^3 x 28.6 nm"
    else:
        N = 10000 if name == "c3" else 400000
        R, sweeps = (1024, 25) if name == "c3" else (148, 3)
        Rc, nx, W = bench.workload_params(N)
        if name == "c4":
            nx = 65
        if name == "c4fine":
            nx = 130
        r = paths.confined_gaussian_walk(N, np.full(N - 1, 16.5), "Spherical", Rc, rng, replicas=R)
        confine = "Spherical"
        if name == "c3":
            hp1 = dict(bench.HP1, cross_talk={"PRC1": -1.0})
            binders, tracks = [hp1, dict(PRC1)], ("H3K9me3", "H3K27me3")
            mu = np.stack([np.linspace(-2.0, 0.0, R), np.linspace(0.0, -2.0, R)], axis=1)
            label = "C3: chromatin 10,000 beads, HP1 + PRC1 (cross-talk -1), mu swept over linspace(-2, 0) across replicas"
        else:
            binders, tracks, mu = [dict(bench.HP1)], ("H3K9me3",), [-1.2]
            label = f"C4: chromatin 400,000 beads, HP1, {nx}^3 voxels, one replica per SM"
    t3, t2 = paths.estimate_tangents_from_coordinates(r)
    nb = len(binders)
    mods = (np.zeros((R, N, nb), dtype=np.int64) if tracks is None
            else chemical_mods.first_beads(N, tracks, replicas=R))
    grid = dict(x_width=W, nx=nx, y_width=W, ny=nx, z_width=W, nz=nx, confine_type=confine, confine_length=Rc,
                vf_limit=0.5)
    ens = ReplicaEnsemble(r, t3, t2, np.zeros((R, N, nb), dtype=np.int64), mods, binders=binders,
                          bond_params=bench.bond_params(N), grid=grid, bead_vol=(4 / 3) * math.pi * 125.0, chi=1.0, mu=mu,
                          moves=default_moves(R, N, 16.5))
    return ens, dict(label=label, replicas=R, beads=N, grid=nx, sweeps_per_step=sweeps, binders=[b["name"] for b in binders])


def run(name, steps, warm):
    ens, cfg = build(name)
    eng = ens.engine
    S = cfg["sweeps_per_step"]
    stream = torch.cuda.ExternalStream(eng.stream())
    for w in range(warm):
        ens.mc_sim(S, 1.0, 10 + w, sync_host=False)
    eng.sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ab, at = [], []
    ev[0].record(stream)
    for k in range(steps):
        ens.mc_sim(S, 1.0, 100 + k, sync_host=False)
        ev[k + 1].record(stream)
        ab.append(eng.last_algo_bytes())
        at.append(eng.last_attempts())
    eng.sync()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[steps])
    kms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    ens.sync()
    peak = 6650.0
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    achieved = sum(ab) / (sum(kms) * 1e-3) / 1e9
    line = dict(metric=bench.METRIC, value=sum(at) / (ms * 1e-3), unit=bench.UNIT, n_gpus=1, steps=steps, warmup=warm,
                ms_per_step=ms / steps, config=cfg,
                launch=dict(warps_per_replica=eng.set_warps_per_replica(0), replicas_per_block=eng.set_replicas_per_block(0),
                            table_slots=eng.set_table_capacity(0)),
                roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                              bytes_per_attempt=sum(ab) / max(1, sum(at))),
                us_per_attempt_and_replica=1e3 * ms / steps / (S * bench.ATTEMPTS_PER_SWEEP),
                acceptance={k: round(float(v), 4) for k, v in ens.acceptance().items()},
                amp_bead_mean=[round(float(x), 2) for x in ens.moves["amp_bead"].mean(axis=0)], hbm_bytes=eng.bytes())
    ens.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="+", choices=["c1", "c3", "c4", "c4fine"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out-dir", default=None)
    ap.add_argument("--lib", default=None, help="another build of the C-ABI library (tools/build_variant.sh), for A/B runs")
    a = ap.parse_args()
    if a.lib:
        sys.path.insert(0, str(ROOT / "tests"))
        import devlib
        devlib.use_library(a.lib)
    for name in a.configs:
        line = run(name, a.steps, a.warmup)
        s = json.dumps(line)
        print(s, flush=True)
        if a.out_dir:
            (Path(a.out_dir) / f"r02_bench_{name}.json").write_text(s + "\n")


if __name__ == "__main__":
    main()
