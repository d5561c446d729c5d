#!/usr/bin/env python
"""BASELINE config 5: parallel-tempering chi ladder, replicas sharded over the GPUs of one box,
NCCL all-gather of one fp64 observable per replica every K sweeps (chromo_b200.parallel).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/c5_exchange.py --replicas 4096 --beads 10000 --rounds 5 --sweeps 10 [--out file.json]

Rank 0 prints one JSON line: attempts/s including the exchange steps, the time of the exchange step alone
(device observable + NCCL all-gather + label update), swap acceptance, and a check that every rank holds the
same ladder (the labels are a permutation of the input ladder)."""
import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=4096, help="total over all ranks")
    ap.add_argument("--beads", type=int, default=10000)
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--sweeps", type=int, default=10, help="MC sweeps between exchange rounds")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import bench
    import oracle as O
    from chromo_b200.ensemble import ReplicaEnsemble, default_moves
    from chromo_b200.parallel import ReplicaExchange, shard_indices
    from chromo_b200.util import poly_paths as paths
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mine = shard_indices(a.replicas, rank, world)
    R, N = len(mine), a.beads
    rng = np.random.default_rng(1000 + rank)
    Rc, nx, W = bench.workload_params(N)
    r = paths.confined_gaussian_walk(N, np.full(N - 1, 16.5), "Spherical", Rc, rng, replicas=R)
    t3, t2 = paths.estimate_tangents_from_coordinates(r)
    mods = paths.synthetic_marks(N, 1, rng, replicas=R)
    hp1 = dict(O.HP1)
    hp1["chemical_potential"] = -1.2
    g = dict(x_width=W, nx=nx, y_width=W, ny=nx, z_width=W, nz=nx, confine_type="Spherical", confine_length=Rc,
             vf_limit=0.5)
    ens = ReplicaEnsemble(r, t3, t2, np.zeros((R, N, 1), dtype=np.int64), mods, binders=[hp1],
                          bond_params=bench.bond_params(N), grid=g, bead_vol=(4 / 3) * math.pi * 125.0,
                          moves=default_moves(R, N, 16.5), device=local)
    ladder = np.geomspace(0.25, 4.0, a.replicas)
    ex = ReplicaExchange(ens, ladder, n_total=a.replicas, seed=3, device=torch.device("cuda", local))
    ens.mc_sim(a.sweeps, 1.0, 7, sync_host=False)  # warm-up (controllers, NCCL communicator)
    ex.step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    t_ex, swaps = 0.0, 0
    for rnd in range(a.rounds):
        ens.mc_sim(a.sweeps, 1.0, 100 + rnd, sync_host=False)
        ens.engine.sync()
        t1 = time.perf_counter()
        swaps += ex.step()
        t_ex += time.perf_counter() - t1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    tt = torch.tensor([wall, t_ex], dtype=torch.float64, device="cuda")
    chk = torch.tensor(ex.chi, dtype=torch.float64, device="cuda")
    same = True
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ref = chk.clone()
        dist.broadcast(ref, 0)
        ok = torch.tensor([float(torch.equal(ref, chk))], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item() == 1.0)
    wall, t_ex = float(tt[0]), float(tt[1])
    attempts = a.replicas * 161 * a.sweeps * a.rounds
    if rank == 0:
        line = dict(config="C5: parallel-tempering chi ladder (geometric 0.25..4), HP1 chromatin", n_gpus=world,
                    replicas=a.replicas, beads=N, sweeps_between_exchanges=a.sweeps, rounds=a.rounds,
                    attempts_per_s=attempts / wall, ms_per_round=1e3 * wall / a.rounds,
                    exchange_ms_per_round=1e3 * t_ex / a.rounds, swaps_accepted=swaps,
                    swap_pairs_tried=a.rounds * (a.replicas // 2),
                    ladder_is_permutation=bool(np.array_equal(np.sort(ex.chi), ladder)),
                    ladders_identical_on_all_ranks=same, backend="nccl" if world > 1 else "single")
        s = json.dumps(line)
        print(s)
        if a.out:
            Path(a.out).write_text(s + "\n")
    ens.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
