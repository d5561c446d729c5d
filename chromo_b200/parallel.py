"""Multi-GPU: replica sharding and the replica-exchange step.

The reference has no distributed code: simulations are
independent.  Here replicas are the unit of sharding -- replica i of a job lives
on rank i % world (one process per GPU, `torch.distributed`), there is no halo
and NO data-path collective in `mc_sim`.  The only exchange is the optional
parallel-tempering step of a chi ladder (BASELINE config 5): every K sweeps each
rank contributes one fp64 observable per local replica, an all-gather (NCCL
over NVLink on GPUs; gloo in the CPU tests) makes the 4,096 x 8 B table visible
everywhere, every rank evaluates the SAME neighbour-swap Metropolis with a
shared counter-based RNG, and the chi LABELS move -- never configurations.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def shard_indices(n_total: int, rank: int, world: int) -> np.ndarray:
    """Global replica ids owned by `rank`: i % world == rank."""
    return np.arange(rank, n_total, world, dtype=np.int64)


def shard_sizes(n_total: int, world: int) -> np.ndarray:
    return np.array([len(range(r, n_total, world)) for r in range(world)], dtype=np.int64)


def _uniforms(seed: int, round_index: int, n: int) -> np.ndarray:
    """Shared uniforms for one exchange round: identical on every rank."""
    return np.random.Generator(np.random.Philox(key=[seed & 0xFFFFFFFFFFFFFFFF, round_index])).random(n)


def swap_decisions(chi: np.ndarray, phi: np.ndarray, round_index: int, seed: int) -> np.ndarray:
    """New chi assignment after one even/odd neighbour-exchange round.

    chi[g], phi[g] are indexed by GLOBAL replica id.  The Hamiltonian sampled by
    the moves is H = H0 + chi * Phi with Phi = sum_bins (V/v) phi^2 (the dE
    convention, fields.pyx:1829-1840), so swapping the labels of replicas a, b
    changes the energy by (chi_a - chi_b) (Phi_b - Phi_a).  Pairs are
    neighbours in the chi-sorted ladder: (0,1),(2,3).. on even rounds,
    (1,2),(3,4).. on odd rounds.
    """
    chi = np.asarray(chi, dtype=float).copy()
    phi = np.asarray(phi, dtype=float)
    order = np.argsort(chi, kind="stable")  # ladder position -> replica id
    start = round_index % 2
    a, b = order[start:len(chi) - 1:2], order[start + 1::2]  # disjoint pairs: decided all at once
    b = b[:len(a)]
    u = _uniforms(seed, round_index, len(a))
    dE = (chi[a] - chi[b]) * (phi[b] - phi[a])
    with np.errstate(over="ignore"):
        swap = u < np.exp(-dE)
    chi[a[swap]], chi[b[swap]] = chi[b[swap]], chi[a[swap]]
    return chi


def all_gather_by_replica(local: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """Gather one fp64 value per replica from every rank into global-id order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local, dtype=float).copy()
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = shard_sizes(n_total, world)
    pad = int(sizes.max())
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    mine = torch.zeros(pad, dtype=torch.float64, device=dev)
    mine[: len(local)] = torch.as_tensor(np.asarray(local, dtype=float), device=dev)
    out = torch.empty(world * pad, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, mine)
    out = out.cpu().numpy().reshape(world, pad)
    glob = np.empty(n_total)
    for r in range(world):
        glob[shard_indices(n_total, r, world)] = out[r, : sizes[r]]
    return glob


class ReplicaExchange:
    """Parallel tempering over a chi ladder for a sharded ensemble."""

    def __init__(self, ensemble, chi_ladder, n_total: Optional[int] = None, seed: int = 0, device=None):
        import torch.distributed as dist
        self.ens = ensemble
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.n_total = int(n_total if n_total is not None else len(chi_ladder))
        self.mine = shard_indices(self.n_total, self.rank, self.world)
        self.chi = np.asarray(chi_ladder, dtype=float).copy()  # by global replica id
        self.seed, self.round, self.device = seed, 0, device
        self.accepted = 0
        self.ens.set_params(chi=self.chi[self.mine])

    def step(self):
        """One exchange round; returns the number of label swaps accepted."""
        phi_local = self.ens.engine.chi_observable()
        phi = all_gather_by_replica(phi_local, self.n_total, self.device)
        new = swap_decisions(self.chi, phi, self.round, self.seed)
        swaps = int(np.count_nonzero(new != self.chi) // 2)
        self.chi = new
        self.ens.set_params(chi=self.chi[self.mine])
        self.round += 1
        self.accepted += swaps
        return swaps
