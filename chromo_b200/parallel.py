"""Multi-GPU: replica sharding and the replica-exchange step.

The reference has no distributed code: simulations are
independent.  Here replicas are the unit of sharding -- rank r of a job owns the
contiguous block of global replica ids [r * R, (r + 1) * R) (one process per GPU,
`torch.distributed`), there is no halo and NO data-path collective in `mc_sim`.
The only exchange is the optional parallel-tempering step of a chi ladder
(BASELINE config 5): every K sweeps each rank computes one fp64 observable per
local replica ON THE DEVICE, one all-gather (NCCL over NVLink; gloo in the CPU
tests) makes the n_total x 8 B table visible everywhere, and every rank runs the
same swap kernel -- a counter-based uniform per pair and round -- which permutes
the chi LABELS; configurations never move.  Observable, all-gather and swap
kernel are queued on the context's stream: no host round trip
(`chromo_exchange_observable` / `chromo_exchange_step`, include/chromo_b200.h).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib


def shard_indices(n_total: int, rank: int, world: int) -> np.ndarray:
    """Global replica ids owned by `rank`: a contiguous block (the first n_total % world ranks hold one more)."""
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return np.arange(start, start + base + (1 if rank < extra else 0), dtype=np.int64)


def _parse_cpulist(text: str) -> list:
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device: int, sysfs: str = "/sys", bdf: str = None) -> dict:
    """Pin the calling process to the host cores of the NUMA node `device` hangs off, so that the pinned host
    buffers allocated afterwards (first touch) and the threads that feed the copy engines are local to the GPU's
    PCIe root: with one rank per GPU on a two-socket host the end-to-end path otherwise crosses the socket link
    for half the ranks.  Call it before allocating pinned memory.  Returns what it did ({"node": n, "cpus": k}, or
    {"node": None, "why": ...} when the platform does not say -- a VM without NUMA information, one node only)."""
    import os
    try:
        if bdf is None:  # the GPU's PCI address
            import torch
            p = torch.cuda.get_device_properties(device)
            bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"{sysfs}/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return dict(node=None, why="the platform reports no NUMA node for the GPU")
        with open(f"{sysfs}/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return dict(node=node, why="none of the node's cores is in this process's cpuset")
        if len(allowed) == len(os.sched_getaffinity(0)):
            return dict(node=node, cpus=len(allowed), why="already local (one node)")
        os.sched_setaffinity(0, allowed)
        return dict(node=node, cpus=len(allowed))
    except Exception as e:  # plumbing, never fatal
        return dict(node=None, why=f"{type(e).__name__}: {e}")


def shard_sizes(n_total: int, world: int) -> np.ndarray:
    return np.array([len(shard_indices(n_total, r, world)) for r in range(world)], dtype=np.int64)


# ---- the swap rule on the host (checker of the device kernel; same counter-based uniforms) -----------------
def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint32 arrays (Salmon et al., SC'11); returns word 0."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m32
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m32
        c1, c3, c0, c2 = p1 & m32, p0 & m32, n0, n2
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return c0


def _uniforms(seed: int, round_index: int, n: int) -> np.ndarray:
    """Uniform of pair p in round t: word 0 of Philox4x32-10(counter = (p, t_lo, t_hi, 'EXCH'), key = seed),
    halved and divided by 2^31 - 1 -- what `exchange_kernel` draws (csrc/field_kernels.cuh)."""
    seed &= 0xFFFFFFFFFFFFFFFF
    p = np.arange(n, dtype=np.uint64)
    w = _philox4x32_10(p, np.full(n, round_index & 0xFFFFFFFF), np.full(n, (round_index >> 32) & 0xFFFFFFFF),
                       np.full(n, 0x45584348), seed & 0xFFFFFFFF, seed >> 32)
    return (w >> np.uint64(1)).astype(np.float64) / 2147483647.0


def swap_decisions(chi: np.ndarray, phi: np.ndarray, round_index: int, seed: int, ladder_len: int = 0) -> np.ndarray:
    """New chi assignment after one even/odd neighbour-exchange round.

    chi[g], phi[g] are indexed by GLOBAL replica id.  The Hamiltonian sampled by
    the moves is H = H0 + chi * Phi with Phi = sum_bins (V/v) phi^2 (the dE
    convention, fields.pyx:1829-1840), so swapping the labels of replicas a, b
    changes the energy by (chi_a - chi_b) (Phi_b - Phi_a).  Pairs are
    neighbours in the chi-sorted ladder: rungs (0,1),(2,3).. on even rounds,
    (1,2),(3,4).. on odd rounds; with `ladder_len` the replicas form independent
    ladders of that many consecutive ids and no pair crosses a boundary.
    """
    chi = np.asarray(chi, dtype=float).copy()
    phi = np.asarray(phi, dtype=float)
    n = len(chi)
    L = n if ladder_len <= 0 else ladder_len
    order = np.concatenate([l0 + np.argsort(chi[l0:l0 + L], kind="stable") for l0 in range(0, n, L)]) if n else \
        np.zeros(0, dtype=np.int64)  # rung -> replica id
    k = np.arange(round_index % 2, n - 1, 2)
    u = _uniforms(seed, round_index, (n + 1) // 2)[: len(k)]  # pair p = (k - start) / 2
    keep = (k + 1) % L != 0
    a, b = order[k], order[k + 1]
    dE = (chi[a] - chi[b]) * (phi[b] - phi[a])
    with np.errstate(over="ignore"):
        swap = keep & (u < np.exp(-dE))
    chi[a[swap]], chi[b[swap]] = chi[b[swap]], chi[a[swap]]
    return chi


def all_gather_by_replica(local: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """Gather one fp64 value per replica from every rank into global-id order (host arrays)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local, dtype=float).copy()
    world = dist.get_world_size()
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
    """Parallel tempering over chi for a sharded ensemble, decided on the device.

    `chi_ladder[g]` is the initial chi of GLOBAL replica g (n_total entries, the same on every rank); this
    rank's ensemble holds replicas [rank * R, (rank + 1) * R).  `ladder_len`: replicas form independent
    ladders of that many consecutive ids (0 = one ladder).  `step()` queues one round on the engine's stream
    and returns at once; `state()` synchronises and reads the permutation back."""

    def __init__(self, ensemble, chi_ladder, n_total: Optional[int] = None, seed: int = 0, device=None,
                 ladder_len: int = 0):
        import torch
        import torch.distributed as dist
        self.ens, self.eng = ensemble, ensemble.engine
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.n_total = int(n_total if n_total is not None else len(chi_ladder))
        R = ensemble.R
        if self.n_total != R * self.world:
            raise ValueError(f"replica exchange needs equal shards: {self.n_total} replicas over {self.world} ranks "
                             f"of {R}")
        self.first = self.rank * R
        self.mine = np.arange(self.first, self.first + R)
        chi = np.ascontiguousarray(chi_ladder, dtype=np.float64)
        if chi.shape != (self.n_total,):
            raise ValueError("chi_ladder must hold one chi per global replica")
        self.ladder_len = int(ladder_len)
        self.seed, self.round = int(seed), 0
        self._L = _lib.lib()
        _lib.check(self._L.chromo_exchange_init(self.eng._h, _lib.dptr(chi), self.n_total, self.first, self.ladder_len))
        self.ens.chi = chi[self.mine].copy()
        # device buffers of the round: this rank's observables and everybody's
        if device is None:
            device = torch.device("cuda", ensemble.device) if torch.cuda.is_available() else torch.device("cpu")
        self.device = torch.device(device)
        self.phi_local = torch.zeros(R, dtype=torch.float64, device=self.device)
        self.phi_all = self.phi_local if self.world == 1 else torch.zeros(self.n_total, dtype=torch.float64,
                                                                           device=self.device)
        self._stream = None
        if self.device.type == "cuda":
            self._stream = torch.cuda.ExternalStream(self.eng.stream(), device=self.device)

    def _queue_round(self):
        import torch.distributed as dist
        L, h = self._L, self.eng._h
        _lib.check(L.chromo_exchange_observable(h, C.c_void_p(self.phi_local.data_ptr())))
        if self.world > 1:  # ordered after the observable kernel, and the swap kernel after it, on the same stream
            dist.all_gather_into_tensor(self.phi_all, self.phi_local)
        _lib.check(L.chromo_exchange_step(h, C.c_void_p(self.phi_all.data_ptr()), self.round, self.seed))
        self.round += 1

    def step(self):
        """Queue one exchange round (observable -> all-gather -> swap kernel) on the engine's stream."""
        if self._stream is not None:
            import torch
            with torch.cuda.stream(self._stream):
                self._queue_round()
        else:
            self._queue_round()

    def state(self):
        """Synchronise and return (rung -> global replica permutation, this rank's chi, pairs tried, swaps accepted)."""
        rung = np.zeros(self.n_total, dtype=np.int32)
        chi = np.zeros(self.ens.R)
        tried, acc = C.c_uint64(0), C.c_uint64(0)
        _lib.check(self._L.chromo_exchange_state(self.eng._h, rung.ctypes.data_as(C.POINTER(C.c_int32)), _lib.dptr(chi),
                                                 C.byref(tried), C.byref(acc)))
        self.ens.chi = chi.copy()
        return rung, chi, int(tried.value), int(acc.value)
