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


def _swap_loop(chi, phi, round_index, seed):
    """Pair-by-pair statement of the exchange rule (the checker of the vectorised form)."""
    chi = np.asarray(chi, dtype=float).copy()
    order = np.argsort(chi, kind="stable")
    pairs = [(order[p], order[p + 1]) for p in range(round_index % 2, len(chi) - 1, 2)]
    u = par._uniforms(seed, round_index, len(pairs))
    for (a, b), ui in zip(pairs, u):
        with np.errstate(over="ignore"):
            if ui < np.exp(-(chi[a] - chi[b]) * (phi[b] - phi[a])):
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
