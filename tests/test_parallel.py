"""Replica sharding and the replica-exchange decision logic, world_size 2 over
gloo on CPU (the N > 1 host path; the GPU data path has no collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chromo_b200 import parallel as par


def test_sharding_partitions_all_replicas():
    for n, w in ((1024, 8), (10, 4), (7, 2), (3, 8)):
        got = np.sort(np.concatenate([par.shard_indices(n, r, w) for r in range(w)]))
        assert np.array_equal(got, np.arange(n))
        assert par.shard_sizes(n, w).sum() == n


def test_numa_binding_reads_the_gpus_node(tmp_path):
    """bind_to_gpu_numa_node against a fake sysfs tree: it takes the cores of the GPU's node that this process may
    use, leaves the process alone when the platform says nothing, and never raises."""
    mine = sorted(os.sched_getaffinity(0))
    dev = tmp_path / "bus/pci/devices/0000:1b:00.0"
    dev.mkdir(parents=True)
    (tmp_path / "devices/system/node/node1").mkdir(parents=True)
    try:
        (dev / "numa_node").write_text("-1\n")
        assert par.bind_to_gpu_numa_node(0, str(tmp_path), "0000:1b:00.0")["node"] is None
        (dev / "numa_node").write_text("1\n")
        (tmp_path / "devices/system/node/node1/cpulist").write_text(f"{mine[0]}\n")
        got = par.bind_to_gpu_numa_node(0, str(tmp_path), "0000:1b:00.0")
        assert got["node"] == 1 and got["cpus"] == 1
        if len(mine) > 1:
            assert sorted(os.sched_getaffinity(0)) == [mine[0]]
        (tmp_path / "devices/system/node/node1/cpulist").write_text("100000-100003\n")  # cores we may not use
        assert "cpuset" in par.bind_to_gpu_numa_node(0, str(tmp_path), "0000:1b:00.0")["why"]
        assert par.bind_to_gpu_numa_node(0, str(tmp_path / "nowhere"), "0000:1b:00.0")["node"] is None
        assert par._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    finally:
        os.sched_setaffinity(0, mine)


def test_swap_rule():
    chi = np.array([0.0, 1.0, 2.0, 3.0])
    # energy-lowering swaps are always accepted: pair (0,1) with Phi_1 < Phi_0
    phi = np.array([5.0, 1.0, 1.0, 5.0])
    new = par.swap_decisions(chi, phi, 0, seed=3)
    # pair (0,1): dE = (0-1)(1-5) = 4 > 0 -> mostly rejected; pair (2,3): dE = (2-3)(5-1) = -4 -> accepted
    assert new[2] == 3.0 and new[3] == 2.0
    # labels are a permutation, and odd rounds pair (1,2)
    assert sorted(new) == sorted(chi)
    new2 = par.swap_decisions(chi, np.array([0.0, 9.0, 0.0, 0.0]), 1, seed=3)
    assert new2[0] == 0.0 and new2[3] == 3.0 and sorted(new2) == sorted(chi)
    # detailed balance: acceptance frequency of an uphill swap ~ exp(-dE)
    acc = np.mean([par.swap_decisions(np.array([0.0, 1.0]), np.array([1.0, 0.0]), 2 * i, seed=i)[0] == 1.0
                   for i in range(4000)])
    assert abs(acc - np.exp(-1.0)) < 0.03


def _swap_loop(chi, phi, round_index, seed, ladder_len=0):
    """Pair-by-pair statement of the exchange rule (the checker of the vectorised form and of the kernel)."""
    chi = np.asarray(chi, dtype=float).copy()
    n = len(chi)
    L = n if ladder_len <= 0 else ladder_len
    order = np.concatenate([l0 + np.argsort(chi[l0:l0 + L], kind="stable") for l0 in range(0, n, L)])
    u = par._uniforms(seed, round_index, (n + 1) // 2)
    for p, k in enumerate(range(round_index % 2, n - 1, 2)):
        if (k + 1) % L == 0:
            continue
        a, b = order[k], order[k + 1]
        with np.errstate(over="ignore"):
            if u[p] < np.exp(-(chi[a] - chi[b]) * (phi[b] - phi[a])):
                chi[a], chi[b] = chi[b], chi[a]
    return chi


def test_vectorised_swaps_equal_the_pairwise_rule():
    rng = np.random.default_rng(0)
    for n in (2, 3, 7, 64, 4096):
        chi = rng.permutation(np.geomspace(0.25, 4.0, n))
        for rnd in range(4):
            phi = rng.normal(size=n) * 3
            want = _swap_loop(chi, phi, rnd, seed=5)
            got = par.swap_decisions(chi, phi, rnd, seed=5)
            assert np.array_equal(got, want), (n, rnd)
            if n % 4 == 0 and n > 4:  # independent ladders of n / 4 rungs
                assert np.array_equal(par.swap_decisions(chi, phi, rnd, 5, n // 4), _swap_loop(chi, phi, rnd, 5, n // 4))
            chi = got
    assert np.array_equal(par.swap_decisions(np.array([1.0]), np.array([0.0]), 0, 1), [1.0])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = par.shard_indices(n_total, rank, world)
        rng = np.random.default_rng(100 + rank)
        chi = np.linspace(0.0, 2.0, n_total)
        hist = []
        for rnd in range(6):
            phi_local = rng.normal(size=len(mine)) + 0.1 * mine
            phi = par.all_gather_by_replica(phi_local, n_total, device="cpu")
            # every rank must see every other rank's values in global-id order
            assert np.allclose(phi[mine], phi_local)
            chi = par.swap_decisions(chi, phi, rnd, seed=11)
            hist.append(chi.copy())
        q.put((rank, np.stack(hist)))
    finally:
        dist.destroy_process_group()


def test_exchange_is_identical_on_all_ranks():
    world, n_total = 2, 13
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0], res[1])                     # same decisions everywhere
    assert np.array_equal(np.sort(res[0][-1]), np.linspace(0.0, 2.0, n_total))  # labels permuted
    assert not np.array_equal(res[0][-1], np.linspace(0.0, 2.0, n_total))       # and something moved


# ---- the device path (chromo_exchange_*), world_size 2 over gloo, kernels on the CPU emulation ------------------
def _device_worker(rank, world, port, q, emu_lib):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        from chromo_b200 import _lib
        from chromo_b200.ensemble import ReplicaEnsemble, default_moves
        import devlib
        devlib.use_library(emu_lib)
        R, N, L = 4, 60, 4  # 4 replicas per rank, two ladders of 4 rungs
        n_total = R * world
        specs = [O.make_spec(N=N, nb=1, seed=70 + rank * R + i) for i in range(R)]
        st = lambda k: np.stack([s[k] for s in specs])
        ens = ReplicaEnsemble(st("r"), st("t3"), st("t2"), st("states"), st("mods"), binders=specs[0]["binders"],
                              bond_params=O.bond_params(specs[0]["bead_length"], 53.0), grid=specs[0]["field"],
                              bead_vol=(4 / 3) * np.pi * 125.0, moves=default_moves(R, N, 16.5),
                              replica_offset=rank * R)
        ladder = np.tile(np.array([0.5, 1.0, 2.0, 4.0]), n_total // L)[np.random.default_rng(1).permutation(n_total)]
        ladder = np.concatenate([np.random.default_rng(2 + l).permutation([0.5, 1.0, 2.0, 4.0]) for l in range(n_total // L)])
        ex = par.ReplicaExchange(ens, ladder, n_total=n_total, seed=17, device="cpu", ladder_len=L)
        chi = ladder.copy()
        hist = []
        for rnd in range(6):
            ens.mc_sim(1, 1.0, 100 + rnd, sync_host=False)
            ex.step()
            ens.engine.sync()
            phi = ex.phi_all.numpy().copy()  # what every rank decided on
            assert np.allclose(phi[ex.mine], ens.engine.chi_observable(), rtol=1e-12)
            chi = par.swap_decisions(chi, phi, rnd, 17, L)  # host statement of the same round
            rung, chi_local, tried, acc = ex.state()
            assert np.array_equal(chi_local, chi[ex.mine]), (rank, rnd)
            hist.append((phi, rung.copy(), chi.copy(), tried, acc))
        # the kernel's chi is what the next mc_sim uses
        assert np.array_equal(ens.chi, chi[ex.mine])
        q.put((rank, hist))
        ens.close()
    finally:
        dist.destroy_process_group()


def test_device_exchange_world2(emu_lib):
    """chromo_exchange_observable -> all-gather -> chromo_exchange_step on two ranks: every rank holds the same
    permutation, it equals the host statement of the rule, labels stay a permutation of the ladders."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_device_worker, args=(r, world, port, q, str(emu_lib))) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    moved = 0
    for (phi0, rung0, chi0, t0, a0), (phi1, rung1, chi1, t1, a1) in zip(res[0], res[1]):
        assert np.array_equal(phi0, phi1) and np.array_equal(rung0, rung1) and np.array_equal(chi0, chi1)
        assert (t0, a0) == (t1, a1)
        for l0 in range(0, 8, 4):  # each ladder keeps its own rungs and replicas
            assert sorted(chi0[l0:l0 + 4]) == [0.5, 1.0, 2.0, 4.0]
            assert sorted(rung0[l0:l0 + 4]) == list(range(l0, l0 + 4))
        moved = a0
    assert res[0][-1][3] == 3 * 4 + 3 * 2  # pairs tried: even rounds 2 per ladder, odd rounds 1 per ladder
    assert moved > 0
