#!/usr/bin/env python
"""bench.py -- MC move attempts/s of the hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): chromatin, 10,000 beads
per replica, HP1 on the first 10,000 entries of the reference's H3K9me3 track
(chromo/chemical_mods/HNCFF683HCZ_H3K9me3_methyl.txt), chi = 1, mu = -1.2,
spherical confinement, 21^3 voxel grid, canonical 161-attempt sweep (30
crank-shaft, 1 end-pivot, 60 slide, 60 tangent-rotation, 10 binding) with
SimpleControl, and 1,024 replicas PER GPU (weak scaling: replicas shard across
ranks, no data-path collective).  One bench "step" = one `mc_sim` call of
--sweeps MC sweeps (default 50; the reference's own scripts call mc_sim with
1,000-6,000 sweeps per snapshot) over every replica = R * sweeps * 161 attempts.

Equal work: SimpleControl keeps widening the bead windows for ~5,000 sweeps until
crank-shaft, end-pivot and slide sit at their upper bound of 150 beads and
tangent rotation at ~14 (profiles/stationary_amplitudes.json); throughput at that
state is half of what it is 500 sweeps in.  BOTH arms therefore START from that
state (controllers on), print the bead windows they ended with, and our arm's
line fails (exit code 3, line still printed) if the CPU leg and the GPU differ by
more than 2 %.

value : attempts/s with the state resident in HBM (CUDA events on the kernel's
        stream around K back-to-back mc_sim launches, max over ranks).
e2e   : the same metric through the host-facing call (ReplicaEnsemble.mc_sim with
        sync_host=True -> chromo_mc_sim_host): per step the r/t3/t2/states/marks
        arrays go pinned-host -> device, the kernel runs, and r/t3/t2/states come
        back -- like handing the reference its numpy arrays.  Copies and kernel are
        pipelined over replica chunks inside the one call.
roofline : dominant kernel = mc_sim_kernel; achieved = algorithmic bytes
        (counted in-kernel with SURVEY.md 8d's per-attempt formula) / kernel time.
cpu_baseline : the reference's own Cython mc_sim (oracle/_ref, kind "reference";
        or the C port of it when that build is absent), one process per host
        core, each an independent replica of the same config, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

ATTEMPTS_PER_SWEEP = 161
METRIC = "MC move attempts/sec (field+elastic dE)"
UNIT = "attempts/s"


# ------------------------------------------------------------------ workload
def workload_params(N: int):
    """Geometry of SURVEY.md 8(d): bead density of the 393,216-bead / 900 nm
    nucleus, voxel ~28.6 nm (one_mark_coarse.py:79,93,117-120)."""
    dens = 393216 / (4.0 / 3.0 * math.pi * 900.0 ** 3)
    Rc = (N / dens / (4.0 * math.pi / 3.0)) ** (1.0 / 3.0)
    n_acc = max(int(round(63 * Rc / 900.0)), 2)
    nx = n_acc + 2
    W = 2 * Rc * (1 + 2.0 / n_acc)
    return Rc, nx, W


HP1 = dict(name="HP1", sites_per_bead=2, bind_energy_mod=-0.01, bind_energy_no_mod=1.52,
           interaction_energy=-4.0, chemical_potential=-1.2, interaction_radius=3.0, cross_talk={"PRC1": 0.0})


def stationary_state():
    """Controller state both arms start from: where SimpleControl settles on this workload
    (tools/stationary_state.py, measured over 10,000 sweeps of 1,024 replicas)."""
    st = dict(amp_bead=[150, 150, 150, 14, 1], amp_move=[0.5238, 0.701, 4.1266, 0.4022, 0.0])
    try:
        d = json.loads((ROOT / "profiles" / "stationary_amplitudes.json").read_text())
        st = dict(amp_bead=[int(x) for x in d["amp_bead"]], amp_move=[float(x) for x in d["amp_move"]])
    except Exception:
        pass
    return st


def stationary_moves(R: int, N: int, first: int = 0):
    """[R,5] controller records of replicas first .. first+R at the stationary state (bead windows clipped to
    the chain's bounds).  The stationary ensemble is a spread, not a point: the rotation / translation
    amplitudes of crank-shaft, end-pivot and slide cycle through [lo, hi] (acceptance stays above 0.5 with the
    window pinned at its bound), and the tangent-rotation window hovers between 14 and 15 beads (mean 14.4).
    Replica i gets a deterministic phase of that cycle, the same in both arms."""
    import numpy as np
    from chromo_b200.ensemble import default_moves
    mv = default_moves(R, N, 16.5)
    st = stationary_state()
    idx = first + np.arange(R)
    phase = (idx * 0.6180339887498949) % 1.0
    for i in range(5):
        lo_b, hi_b = int(mv["bead_amp_lo"][0, i]), int(mv["bead_amp_hi"][0, i])
        lo_m, hi_m = float(mv["move_amp_lo"][0, i]), float(mv["move_amp_hi"][0, i])
        mv["amp_bead"][:, i] = max(lo_b, min(int(st["amp_bead"][i]), hi_b))
        mv["amp_move"][:, i] = max(lo_m, min(st["amp_move"][i], hi_m))
        mv["acceptance_rate"][:, i] = 0.5
        if i in (0, 1, 2) and hi_m > lo_m:  # geometric sawtooth lo -> hi (x 1/0.95 per sweep)
            mv["amp_move"][:, i] = lo_m * (hi_m / lo_m) ** phase
    if st["amp_bead"][3] == 14:
        mv["amp_bead"][:, 3] = np.clip(np.where(idx % 5 < 2, 15, 14), int(mv["bead_amp_lo"][0, 3]), int(mv["bead_amp_hi"][0, 3]))
    return mv


def make_inputs(R: int, N: int, seed: int, pinned: bool, nb: int = 1):
    """Synthetic replicas: confined Gaussian-direction walks, central-difference tangents, the first N
    entries of the reference's mark track(s) (the same for every replica), all binding states 0."""
    import numpy as np
    from chromo_b200.util import chemical_mods
    from chromo_b200.util import poly_paths as paths
    rng = np.random.default_rng(seed)
    Rc, nx, W = workload_params(N)
    r = paths.confined_gaussian_walk(N, np.full(N - 1, 16.5), "Spherical", Rc, rng, replicas=R)
    t3, t2 = paths.estimate_tangents_from_coordinates(r)
    mods = chemical_mods.first_beads(N, chemical_mods.TRACKS[:nb], replicas=R)
    states = np.zeros((R, N, nb), dtype=np.int64)
    if pinned:
        import torch

        def pin(a):
            t = torch.empty(a.shape, dtype=torch.float64 if a.dtype == np.float64 else torch.int64, pin_memory=True)
            v = t.numpy()
            v[...] = a
            return v, t
        keep = []
        out = []
        for a in (r, t3, t2, states, mods):
            v, t = pin(np.ascontiguousarray(a))
            keep.append(t)
            out.append(v)
        r, t3, t2, states, mods = out
        make_inputs._keep = keep
    grid = dict(x_width=W, nx=nx, y_width=W, ny=nx, z_width=W, nz=nx, confine_type="Spherical",
                confine_length=Rc, vf_limit=0.5)
    return r, t3, t2, states, mods, grid


def bond_params(N: int, spacing=16.5, lp=53.0, lt=None):
    """SSWLC._find_parameters (polymers.pyx:1545-1601) for uniform spacing; with `lt` also the twist
    parameters of SSTWLC (polymers.pyx:2000, 2088-2090)."""
    import numpy as np
    from chromo_b200.util import dss_params as tab
    d = spacing / lp
    vals = dict(eps_bend=np.interp(d, tab[:, 0], tab[:, 1]) / d,
                gamma=np.interp(d, tab[:, 0], tab[:, 2]) * d * lp,
                eps_par=np.interp(d, tab[:, 0], tab[:, 3]) / (d * lp ** 2),
                eps_perp=np.interp(d, tab[:, 0], tab[:, 4]) / (d * lp ** 2),
                eta=np.interp(d, tab[:, 0], tab[:, 5]) / lp)
    if lt is not None:
        vals["eps_twist"] = lt / (d * lp)
        vals["natural_twist"] = spacing * (2 * np.pi / 10.5) / 0.332
    return {k: np.full(N - 1, v) for k, v in vals.items()}


# ------------------------------------------------------------- clock sampling
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------- CPU baseline
def _ref_worker(args):
    """One host core: the reference's Cython mc_sim on one replica of the workload, started from the
    controllers' stationary state like the GPU arm (a few warm-up sweeps decorrelate the replicas)."""
    rank, N, warm, sweeps, reps, kind = args
    import numpy as np
    sys.path.insert(0, str(ROOT / "oracle"))
    sys.path.insert(0, str(ROOT))
    r, t3, t2, states, mods, grid = make_inputs(1, N, 1000 + rank, pinned=False)
    spec = dict(N=N, nb=1, r=r[0], t3=t3[0], t2=t2[0], states=states[0], mods=mods[0],
                bead_length=np.full(N - 1, 16.5), lp=53.0, bead_rad=5.0, binders=[dict(HP1)], max_binders=-1,
                field=dict(grid, chi=1.0))
    import oracle as O
    times = []
    if kind == "reference":
        poly, df, field, M = O.ref_objects(spec)
        ctrl, mc, mcs, sh = M["mc_controller"], M["mc"], M["mc_sim"], M["shim"]
        bb, mb = mc.get_amplitude_bounds([poly])
        cs = ctrl.all_moves("/tmp/chromo_bench", bb.bounds, mb.bounds, ctrl.SimpleControl)
        st = stationary_moves(1, N, first=rank)
        for i, c in enumerate(cs):
            c.move.amp_bead = int(st["amp_bead"][0, i])
            c.move.amp_move = float(st["amp_move"][0, i])
            c.move.acceptance_tracker.acceptance_rate = 0.5
        sh.c_srand(rank + 1)
        with np.errstate(over="ignore"):
            mcs.mc_sim([poly], df, warm, cs, field, 1.0, rank)
            for k in range(reps):
                t0 = time.perf_counter()
                mcs.mc_sim([poly], df, sweeps, cs, field, 1.0, rank + 7 + k)
                times.append(time.perf_counter() - t0)
        amp = [c.move.amp_bead for c in cs]
    else:
        o = O.OracleSim(spec, srand_seed=rank + 1)
        mv = O.make_moves(N, 16.5)
        st = stationary_moves(1, N, first=rank)
        for i in range(5):
            mv[i].amp_bead, mv[i].amp_move = int(st["amp_bead"][0, i]), float(st["amp_move"][0, i])
            mv[i].acceptance_rate = 0.5
        o.mc_sim(mv, warm, rank)
        for k in range(reps):
            t0 = time.perf_counter()
            o.mc_sim(mv, sweeps, rank + 7 + k)
            times.append(time.perf_counter() - t0)
        amp = [m.amp_bead for m in mv]
    return sweeps * ATTEMPTS_PER_SWEEP, times, amp


def cpu_reference_sample(N: int, sweeps: int, warm: int, reps: int = 1, cores: int | None = None):
    """One process per host core; per repetition: sum of attempts / max wall time."""
    import multiprocessing as mp
    sys.path.insert(0, str(ROOT / "oracle"))
    import oracle as O
    kind = "reference" if O.ref_available() else "port"
    if kind == "port":
        O.build_lib()
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_worker, [(i, N, warm, sweeps, reps, kind) for i in range(cores)], chunksize=1)
    per_rep = []
    for k in range(reps):
        attempts = sum(a for a, _, _ in res)
        tmax = max(t[k] for _, t, _ in res)
        per_rep.append((attempts / tmax, tmax))
    value = sum(v for v, _ in per_rep) / reps
    amp = [sum(a[i] for _, _, a in res) / len(res) for i in range(5)]
    cb = dict(value=value, unit=UNIT, cores=cores, kind=kind,
              sample=f"{cores} processes (one per host core) x {sweeps} MC sweeps ({sweeps * ATTEMPTS_PER_SWEEP} attempts "
                     f"each) of one N={N} HP1 replica, started at SimpleControl's stationary state (bead windows "
                     f"150/150/150/14/1) + {warm} untimed sweeps; wall time of mc_sim only",
              single_core=max(a / min(t) for a, t, _ in res), amp_bead_mean=[round(x, 2) for x in amp])
    return cb, per_rep


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.beads
    cb, per_rep = cpu_reference_sample(N, args.ref_sweeps, args.ref_warm, reps=args.warmup + args.steps)
    timed = per_rep[args.warmup:]
    v = sum(x for x, _ in timed) / len(timed)
    cb["value"] = v
    line = dict(metric=METRIC, value=v, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * sum(t for _, t in timed) / len(timed), higher_is_better=True,
                scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=workload_config(args), cpu_baseline=cb, amp_bead_mean=cb.get("amp_bead_mean"),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


def workload_config(args):
    Rc, nx, W = workload_params(args.beads)
    label = {10000: "C2", 400000: "C4-sized", 1000: "C1-sized"}.get(args.beads, "custom size")
    if args.beads == 10000 and args.replicas != 1024:
        label = "C2-sized"
    twist = f", SSTWLC twist lt={args.lt:g}" if getattr(args, "lt", None) is not None else ""
    state_mb = args.replicas * args.beads * 74 / 1e6   # r, t3, t2 fp64 + states, marks int8
    field_mb = args.replicas * nx ** 3 * 16 / 1e6      # (bead, HP1) fp64 per voxel
    fits = state_mb + field_mb <= 126
    st = stationary_state()
    return dict(workload=f"{label}: chromatin {args.beads} beads, HP1 on H3K9me3 (first {args.beads} entries of the "
                         f"reference's track), chi=1, mu=-1.2, {args.replicas} replicas per GPU, {nx}^3 voxels, "
                         f"spherical confinement R={Rc:.1f} nm, SimpleControl started at its stationary state{twist}",
                replicas_per_gpu=args.replicas, beads=args.beads, grid=nx, sweeps_per_step=args.sweeps,
                attempts_per_sweep=ATTEMPTS_PER_SWEEP, start_amp_bead=st["amp_bead"], start_amp_move=st["amp_move"],
                l2=(f"working set (state {state_mb:.0f} MB + field {field_mb:.0f} MB per GPU) "
                    + ("FITS the 126 MB L2: not a cold-cache number" if fits
                       else "exceeds the 126 MB L2; no flush needed")),
                parallelism=f"replica-sharded x{args.gpus}")


# --------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch
    from chromo_b200.ensemble import ReplicaEnsemble

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: chromo_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    from chromo_b200.parallel import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local) if world > 1 else dict(node=None, why="one rank")  # before the pinned buffers
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    R, N, S, K, Wm = args.replicas, args.beads, args.sweeps, args.steps, max(args.warmup, 3)

    r, t3, t2, states, mods, grid = make_inputs(R, N, 1234 + rank, pinned=True)
    ens = ReplicaEnsemble(r, t3, t2, states, mods, binders=[dict(HP1)], bond_params=bond_params(N, lt=args.lt), grid=grid,
                          bead_vol=(4 / 3) * math.pi * 5.0 ** 3, chi=1.0, mu=[-1.2],
                          moves=stationary_moves(R, N, first=rank * R), device=local,
                          replica_offset=rank * R)  # every rank's replicas draw their own random streams
    eng = ens.engine
    if args.batch:
        eng.set_batch_size(args.batch)
    warps = eng.set_warps_per_replica(args.warps)
    rpb = eng.set_replicas_per_block(args.rpb)
    cap = eng.set_table_capacity(args.table_slots)
    stream = torch.cuda.ExternalStream(eng.stream(), device=local)

    def barrier():
        torch.cuda.synchronize()
        eng.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput --------------------------------------
    # both arms start at SimpleControl's stationary state; the warm-up steps decorrelate the replicas
    for w in range(Wm):
        ens.mc_sim(S, 1.0, 100 + w, sync_host=False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    algo_bytes, attempts = [], []
    ev[0].record(stream)
    for k in range(K):
        ens.mc_sim(S, 1.0, 1000 + k, sync_host=False)
        ev[k + 1].record(stream)
        # per-launch counters (a stream sync + two 8 KB reads: microseconds against a ~100 ms launch)
        algo_bytes.append(eng.last_algo_bytes())
        attempts.append(eng.last_attempts())
    eng.sync()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev[0].elapsed_time(ev[K])
    kernel_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(K)]
    ens.sync()
    acc = ens.acceptance()
    amp_bead_mean = [float(x) for x in ens.moves["amp_bead"].mean(axis=0)]
    hbm_bytes = eng.bytes()
    barrier()

    # ---- end to end through the host-facing call --------------------------
    # per call: r, t3, t2 (fp64) and states (int64) go host -> device and come back; the marks go up once
    # per ensemble (they never change: ReplicaEnsemble.mc_sim keeps them resident)
    h2d = (3 * R * N * 24 + R * N * 1 * 8) * world
    d2h = (3 * R * N * 24 + R * N * 1 * 8) * world
    Ke = max(1, min(K, args.e2e_steps))
    nblk = -(-R // rpb)
    e2e_chunks = 8  # (the library's own rule for n_chunks = 0, repeated here to count launches)
    while e2e_chunks > 1 and e2e_chunks * -(-nblk // e2e_chunks) > torch.cuda.get_device_properties(local).multi_processor_count:
        e2e_chunks -= 1
    ens.mc_sim(S, 1.0, 5000, sync_host=True)  # (refreshes the host arrays from the device first)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for k in range(Ke):
        ens.mc_sim(S, 1.0, 6000 + k, sync_host=True)
    e1.record(stream)
    eng.sync()
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
    barrier()
    # the link alone: every rank uploads / downloads its whole state at the same time (what the e2e path
    # contends for at N > 1); r, t3, t2 as f64, states (and, on upload, marks) as int64
    link = []
    for fn, nbytes in ((ens.push, 3 * R * N * 24 + 2 * R * N * 8), (ens.pull, 3 * R * N * 24 + R * N * 8)):
        eng.sync()
        barrier()
        t0 = time.perf_counter()
        fn()
        eng.sync()
        link.append(nbytes / (time.perf_counter() - t0) / 1e9)
    barrier()

    # ---- extra legs: strong scaling of C2 and the C5 replica-exchange ladder ---------------
    extra = {}
    if not args.no_extra_legs:
        ens.close()
        ens = None
        extra = extra_legs(args, rank, world, local, dist, (r, t3, t2, states, mods, grid))

    # ---- reduce over ranks --------------------------------------------------
    t = torch.tensor([ms_total, e2e_ms, -link[0], -link[1]], dtype=torch.float64, device=f"cuda:{local}")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    link = [-float(t[2]), -float(t[3])]  # the slowest rank's
    attempts_step = R * S * ATTEMPTS_PER_SWEEP * world
    value = attempts_step * K / (ms_total * 1e-3)
    e2e = attempts_step * Ke / (e2e_ms * 1e-3)

    rc = 0
    if rank == 0:
        peaks, peak_src = 6650.0, "fallback"
        try:
            mp = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            peaks, peak_src = float(mp["hbm_gbs"]), "measured"
        except Exception:
            pass
        # roofline of the dominant kernel: algorithmic bytes of ALL timed launches / their summed CUDA-event time
        k_ms = sum(kernel_ms) / len(kernel_ms)
        bytes_launch = sum(algo_bytes) / len(algo_bytes)
        attempts_launch = sum(attempts) / len(attempts)
        achieved = sum(algo_bytes) / (sum(kernel_ms) * 1e-3) / 1e9
        traffic = None
        try:
            tr = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
            per_attempt = tr.get("mc_sim_kernel", {}).get("dram_bytes_per_attempt")
            # ncu dram__bytes_read + dram__bytes_write of one launch at THIS working point, per attempt
            traffic = per_attempt * attempts_launch if per_attempt else None
        except Exception:
            pass
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=Wm,
            ms_per_step=ms_total / K, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
            data="synthetic", config=workload_config(args),
            launch=dict(warps_per_replica=warps, replicas_per_block=rpb, table_slots=cap, rng="philox4x32-10"),
            e2e=dict(value=e2e, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=Ke,
                     ms_per_step=e2e_ms / Ke, replica_chunks=e2e_chunks,
                     h2d_link_gbs_per_gpu=link[0], d2h_link_gbs_per_gpu=link[1],  # whole-state copies alone, all ranks at once, slowest rank
                     host_numa_binding_rank0=numa,
                     path="chromo_mc_sim_host: pinned host arrays -> device -> kernel -> host, pipelined over replica chunks"),
            gpu_launches=K + 3 * e2e_chunks * Ke,  # e2e: per replica chunk 1 narrowing kernel (states; the marks stay resident), the MC kernel, 1 widening
            clocks=clocks,
            roofline=dict(bound="hbm", achieved=achieved, peak=peaks, unit="GB/s", frac=achieved / peaks,
                          traffic=traffic, peak_source=peak_src, kernel=f"mc_sim_kernel<PhiloxRng,1,{warps}>",
                          kernel_ms=k_ms, algorithmic_bytes_per_launch=bytes_launch,
                          attempts_per_launch=attempts_launch,
                          bytes_per_attempt=bytes_launch / max(1.0, attempts_launch),
                          note="achieved = sum of algorithmic bytes of the timed launches / sum of their CUDA-event "
                               "times; the kernel is instruction-supply bound, not bandwidth-bound: DESIGN.md 4.1"),
            acceptance={k: round(float(v), 4) for k, v in acc.items()},
            amp_bead_mean=[round(x, 2) for x in amp_bead_mean],
            hbm_bytes=hbm_bytes,
            **extra,
        )
        if world == 1 and not args.no_cpu_baseline:
            try:
                # the same code path and start state as `--impl reference`: 1 untimed + 12 timed repetitions
                cb, per_rep = cpu_reference_sample(N, args.ref_sweeps, args.ref_warm, reps=13)
                cb["value"] = sum(v for v, _ in per_rep[1:]) / len(per_rep[1:])
                line["cpu_baseline"] = cb
                ours, theirs = line["amp_bead_mean"], cb["amp_bead_mean"]
                # bead windows within 2 %; the tangent-rotation window is an integer hovering between 14 and 15,
                # so a handful of CPU replicas can sit up to one bead from the GPU's 1,024-replica mean
                diff = max(abs(a - b) / max(abs(a), abs(b), 1e-9) for i, (a, b) in enumerate(zip(ours, theirs)) if i != 3)
                tdiff = abs(ours[3] - theirs[3])
                ok = diff <= 0.02 and tdiff <= 1.0
                line["equal_work"] = dict(amp_bead_gpu=ours, amp_bead_cpu=theirs, max_rel_diff=round(diff, 4),
                                          tangent_window_diff_beads=round(tdiff, 2), ok=ok)
                if not ok:
                    rc = 3
                    print(f"bench.py: the arms did NOT do equal work: bead windows {ours} (GPU) vs {theirs} (CPU)",
                          file=sys.stderr)
            except Exception as e:  # never lose the GPU number to a baseline hiccup
                line["cpu_baseline"] = dict(value=None, unit=UNIT, cores=0, kind="unavailable", sample=str(e)[:200])
        emit(line)
    if ens is not None:
        ens.close()
    if dist is not None:
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


def extra_legs(args, rank, world, local, dist, inputs):
    """Two more numbers on the same line (every rank takes part; the values are maxima over ranks):
    strong_scaling  C2 with 1,024 replicas IN TOTAL, split over the ranks (the main value is weak scaling);
    c5_exchange     BASELINE config 5: 4,096 replicas in total, 256 chi ladders of 16 rungs (geometric 0.25 .. 4:
                    ~35 % of the neighbour swaps accepted), one exchange round every 10 sweeps INSIDE the timed
                    region: device observable -> NCCL all-gather (8 B per replica) -> swap kernel, no host trip."""
    import numpy as np
    import torch
    from chromo_b200 import parallel as par
    from chromo_b200.ensemble import ReplicaEnsemble
    N, S = args.beads, args.sweeps
    r, t3, t2, states, mods, grid = inputs
    dev = torch.device("cuda", local)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def rank_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    out = {}
    # ---- strong scaling: the same 1,024 replicas as one GPU runs, R / world per rank ----
    total = args.replicas
    if world > 1 and total % world == 0:
        Rs = total // world
        sl = slice(0, Rs)
        ens = ReplicaEnsemble(r[sl].copy(), t3[sl].copy(), t2[sl].copy(), states[sl].copy(), mods[sl], binders=[dict(HP1)],
                              bond_params=bond_params(N, lt=args.lt), grid=grid, bead_vol=(4 / 3) * math.pi * 5.0 ** 3,
                              chi=1.0, mu=[-1.2], moves=stationary_moves(Rs, N, first=rank * Rs), device=local,
                              replica_offset=rank * Rs)
        stream = torch.cuda.ExternalStream(ens.engine.stream(), device=local)
        for w in range(3):
            ens.mc_sim(S, 1.0, 300 + w, sync_host=False)
        ens.engine.sync()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Ks = 4
        e0.record(stream)
        for k in range(Ks):
            ens.mc_sim(S, 1.0, 400 + k, sync_host=False)
        e1.record(stream)
        ens.engine.sync()
        torch.cuda.synchronize()
        ms = rank_max(e0.elapsed_time(e1))
        out["strong_scaling"] = dict(value=total * S * ATTEMPTS_PER_SWEEP * Ks / (ms * 1e-3), unit=UNIT, replicas_total=total,
                                     replicas_per_gpu=Rs, steps=Ks, ms_per_step=ms / Ks, launch=dict(
                                         warps_per_replica=ens.engine.set_warps_per_replica(0),
                                         replicas_per_block=ens.engine.set_replicas_per_block(0)))
        ens.close()
    # ---- C5: replica exchange ----
    total5, L = 4096, 16
    if total5 % world == 0 and (total5 // world) % L == 0 and N <= 20000:
        R5 = total5 // world
        r5, t35, t25, st5, md5, _ = make_inputs(R5, N, 4321 + rank, pinned=False)
        ladder = np.tile(np.geomspace(0.25, 4.0, L), total5 // L)
        ens = ReplicaEnsemble(r5, t35, t25, st5, md5, binders=[dict(HP1)], bond_params=bond_params(N), grid=grid,
                              bead_vol=(4 / 3) * math.pi * 5.0 ** 3, chi=ladder[rank * R5:(rank + 1) * R5], mu=[-1.2],
                              moves=stationary_moves(R5, N, first=rank * R5), device=local, replica_offset=rank * R5)
        ex = par.ReplicaExchange(ens, ladder, n_total=total5, seed=11, device=dev, ladder_len=L)
        stream = torch.cuda.ExternalStream(ens.engine.stream(), device=local)
        sweeps, rounds = 10, 6
        ens.mc_sim(100, 1.0, 499, sync_host=False)  # the rungs drift apart before the ladder is exercised
        for w in range(4):  # warm-up: kernels, NCCL communicator
            ens.mc_sim(sweeps, 1.0, 500 + w, sync_host=False)
            ex.step()
        _, _, tried0, acc0 = ex.state()
        sync_all()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * rounds + 2)]
        ev[0].record(stream)
        for k in range(rounds):
            ens.mc_sim(sweeps, 1.0, 600 + k, sync_host=False)
            ev[1 + 2 * k].record(stream)
            ex.step()
            ev[2 + 2 * k].record(stream)
        ens.engine.sync()
        torch.cuda.synchronize()
        ms = rank_max(ev[0].elapsed_time(ev[2 * rounds]))
        ex_us = rank_max(1e3 * sum(ev[1 + 2 * k].elapsed_time(ev[2 + 2 * k]) for k in range(rounds)) / rounds)
        rung, chi_local, tried, acc = ex.state()
        phi = ex.phi_all.cpu().numpy().copy()
        # the exchange machinery alone: back-to-back rounds with the ranks in lockstep (inside the run above the
        # all-gather also waits for the slowest rank's sweeps -- that wait is load imbalance, not exchange cost)
        sync_all()
        reps = 20
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record(stream)
        for k in range(reps):
            ex.step()
        x1.record(stream)
        ens.engine.sync()
        torch.cuda.synchronize()
        ex_alone_us = rank_max(1e3 * x0.elapsed_time(x1) / reps)
        same = True
        if dist is not None:  # every rank must hold the same permutation
            a = torch.as_tensor(rung.astype(np.int64), device=dev)
            lo, hi = a.clone(), a.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(lo, hi))
        by_rung = np.stack([phi[rung[l0:l0 + L]] for l0 in range(0, total5, L)]).mean(axis=0)
        out["c5_exchange"] = dict(
            value=total5 * sweeps * ATTEMPTS_PER_SWEEP * rounds / (ms * 1e-3), unit=UNIT, replicas_total=total5,
            replicas_per_gpu=R5, ladders=total5 // L, rungs_per_ladder=L, chi_range=[0.25, 4.0],
            sweeps_between_exchanges=sweeps, rounds=rounds, ms_per_round=ms / rounds,
            exchange_us_per_round=ex_alone_us, exchange_us_per_round_incl_wait_for_slowest_rank=ex_us,
            swap_acceptance=(acc - acc0) / max(1, tried - tried0), ladders_identical_on_all_ranks=same,
            labels_are_permutations=bool(all(sorted(rung[l0:l0 + L]) == list(range(l0, l0 + L)) for l0 in range(0, total5, L))),
            mean_phi_first_rung=float(by_rung[0]), mean_phi_last_rung=float(by_rung[-1]),
            collective="all_gather_into_tensor over NCCL, 8 B per replica, on the kernel stream" if world > 1 else "none (one rank)")
        ens.close()
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """Route file descriptor 1 to stderr while the bench runs (NCCL and friends print banners to
    stdout); `emit` writes the ONE JSON line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=1024, help="replicas per GPU")
    ap.add_argument("--beads", type=int, default=10000)
    ap.add_argument("--sweeps", type=int, default=50, help="MC sweeps (161 attempts each) per bench step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-sweeps", type=int, default=20, help="MC sweeps per process and repetition in the CPU reference sample")
    ap.add_argument("--ref-warm", type=int, default=10, help="untimed sweeps of the CPU reference after the stationary start")
    ap.add_argument("--table-slots", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0, help="warps per replica in the MC kernel (0 = library default)")
    ap.add_argument("--rpb", type=int, default=0, help="replicas per thread block (0 = library default)")
    ap.add_argument("--batch", type=int, default=0, help="attempts prepared at once (1..32; 0 = library default, 32)")
    ap.add_argument("--lt", type=float, default=None, help="twist persistence length: run the SSTWLC kernels "
                    "(not the headline configuration; the reference arm ignores it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the strong-scaling and C5 exchange legs")
    ap.add_argument("--lib", default=None, help="development: load this build of libchromo_b200.so (A/B timing)")
    args = ap.parse_args()
    quiet_stdout()
    if args.lib:
        sys.path.insert(0, str(ROOT / "tests"))
        import devlib
        devlib.use_library(args.lib)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
