"""BASELINE.json's full sizes on the GPU, checked through size-independent
properties (the oracle cannot run these sizes in seconds):
  * mass conservation  sum(rho * V) = N  (validate_density_calculation.ipynb)
  * the incrementally updated density equals a full recompute
  * the oracle agrees on a sampled replica's total energies and bin sets
  * states stay in [0, sites], beads stay inside the confinement, |t3| = 1
  * idempotence: recompute twice -> same field
"""
import math

import numpy as np
import pytest

import oracle as O
from common import close

pytestmark = pytest.mark.gpu


def _ensemble(R, N, nb=1, seed=0, grid=None, chi=1.0, mu=None, cross_talk=0.0):
    import bench
    from chromo_b200.ensemble import ReplicaEnsemble, default_moves
    from chromo_b200.util import poly_paths as paths
    rng = np.random.default_rng(seed)
    Rc, nx, W = bench.workload_params(N)
    nx = grid or nx
    r = paths.confined_gaussian_walk(N, np.full(N - 1, 16.5), "Spherical", Rc, rng, replicas=R)
    t3, t2 = paths.estimate_tangents_from_coordinates(r)
    mods = paths.synthetic_marks(N, nb, rng, replicas=R)
    states = np.zeros((R, N, nb), dtype=np.int64)
    binders = [dict(O.HP1), dict(O.PRC1)][:nb]
    for b in binders:
        b["chemical_potential"] = -1.2
    if nb == 2:
        binders[0]["cross_talk"] = {"PRC1": cross_talk}
    g = dict(x_width=W, nx=nx, y_width=W, ny=nx, z_width=W, nz=nx, confine_type="Spherical",
             confine_length=Rc, vf_limit=0.5)
    ens = ReplicaEnsemble(r, t3, t2, states, mods, binders=binders, bond_params=bench.bond_params(N), grid=g,
                          bead_vol=(4 / 3) * math.pi * 125.0, chi=chi, mu=mu, moves=default_moves(R, N, 16.5))
    return ens, g, binders, Rc


def _spec_of(ens, g, binders, rep, chi=1.0):
    N = ens.N
    return dict(N=N, nb=ens.nb, r=ens.r[rep], t3=ens.t3[rep], t2=ens.t2[rep], states=ens.states[rep],
                mods=ens.chemical_mods[rep], bead_length=np.full(N - 1, 16.5), lp=53.0, bead_rad=5.0,
                binders=binders, max_binders=-1, field=dict(g, chi=chi))


def _invariants(ens, g, Rc, sites=2):
    dens = ens.density()
    vol_bin = g["x_width"] ** 3 / g["nx"] ** 3
    assert np.allclose(dens[..., 0].sum(axis=1) * vol_bin, ens.N, rtol=1e-9)
    ens.engine.field_recompute(clamp=False)
    d2 = ens.density()
    assert np.allclose(dens, d2, rtol=1e-9, atol=1e-12 * d2.max())
    ens.engine.field_recompute(clamp=False)
    assert np.allclose(ens.density(), d2, rtol=1e-13, atol=0)
    ens.pull()
    assert ens.states.min() >= 0 and ens.states.max() <= sites
    assert np.all(np.linalg.norm(ens.r, axis=2) <= Rc + 1e-9)
    assert np.allclose(np.linalg.norm(ens.t3, axis=2), 1.0, atol=1e-9)
    assert np.allclose(np.linalg.norm(ens.t2, axis=2), 1.0, atol=1e-9)
    # per column: density of bound proteins = sum over beads of state * weights -> total = sum(states)/V
    for b in range(ens.nb):
        assert np.allclose(dens[..., 1 + b].sum(axis=1) * vol_bin, ens.states[..., b].sum(axis=1), rtol=1e-9)


def test_c2_full_size(cuda_backend):
    """C2: chromatin 10,000 beads, HP1, chi = 1 (64 replicas here; the bench runs 1,024)."""
    ens, g, binders, Rc = _ensemble(64, 10_000, nb=1, seed=1)
    ens.mc_sim(4, 1.0, 17, sync_host=False)
    ens.sync()
    assert ens.engine.last_attempts() == 64 * 4 * 161
    _invariants(ens, g, Rc)
    # one replica against the oracle: total energies + occupied-bin set of its final state
    ens.pull()
    spec = _spec_of(ens, g, binders, 5)
    o = O.OracleSim(spec)
    d = ens.density()[5]
    assert np.array_equal(d != 0, o.density != 0)
    assert np.allclose(d, o.density, rtol=1e-9, atol=0)
    assert close(ens.field_energy()[5], o.field_E())
    assert close(ens.elastic_energy()[5], o.poly_E())
    acc = ens.acceptance()
    assert all(0.05 < a < 0.99 for a in acc.values()), acc
    ens.close()


def test_c3_two_binders_with_sweep(cuda_backend):
    """C3: HP1 + PRC1 with cross-talk and a chemical-potential sweep across replicas."""
    R = 32
    mu = np.stack([np.linspace(-2.0, 0.0, R), np.linspace(0.0, -2.0, R)], axis=1)
    ens, g, binders, Rc = _ensemble(R, 10_000, nb=2, seed=2, mu=mu, cross_talk=-1.0)
    ens.mc_sim(6, 1.0, 23, sync_host=False)
    ens.sync()
    _invariants(ens, g, Rc)
    ens.pull()
    bound = ens.states.sum(axis=1)  # [R, 2] proteins bound per replica
    # more favourable chemical potential -> more binding (monotone trend across the sweep)
    assert bound[-1, 0] > bound[0, 0] and bound[0, 1] > bound[-1, 1]
    spec = _spec_of(ens, g, binders, 7)
    for b, m in zip(spec["binders"], mu[7]):
        b["chemical_potential"] = float(m)
    o = O.OracleSim(spec)
    assert close(ens.field_energy()[7], o.field_E())
    ens.close()


def test_c4_chromosome_scale(cuda_backend):
    """C4: 400,000 beads, 65^3 grid (2 replicas): HBM-resident field, same invariants."""
    ens, g, binders, Rc = _ensemble(2, 400_000, nb=1, seed=3, grid=65)
    ens.mc_sim(2, 1.0, 29, sync_host=False)
    ens.sync()
    _invariants(ens, g, Rc)
    ens.pull()
    o = O.OracleSim(_spec_of(ens, g, binders, 1))
    assert np.array_equal(ens.density()[1] != 0, o.density != 0)
    assert close(ens.elastic_energy()[1], o.poly_E())
    ens.close()


def test_c4_fine_grid(cuda_backend):
    """C4 on the fine 130^3 grid (2.2 M voxels per replica): nothing in the kernels scans the grid
    per move (the reference clears and walks all of it twice per move, fields.pyx:1223-1226, 1971-1975), so a
    fine grid costs HBM footprint only.  Same invariants, the oracle agrees on the occupied voxels and the total
    energies of one replica, and a short production run equals the oracle's on the same streams."""
    ens, g, binders, Rc = _ensemble(2, 400_000, nb=1, seed=8, grid=130)
    r0, t30, t20, st0 = ens.r.copy(), ens.t3.copy(), ens.t2.copy(), ens.states.copy()
    spec0 = dict(N=400_000, nb=1, r=r0[1], t3=t30[1], t2=t20[1], states=st0[1], mods=ens.chemical_mods[1],
                 bead_length=np.full(399_999, 16.5), lp=53.0, bead_rad=5.0, binders=binders, max_binders=-1,
                 field=dict(g, chi=1.0))
    ens.mc_sim(2, 1.0, 41, sync_host=True)
    _invariants(ens, g, Rc)
    o = O.OracleSim(_spec_of(ens, g, binders, 1))
    d = ens.density()[1]
    assert np.array_equal(d != 0, o.density != 0)
    assert close(ens.field_energy()[1], o.field_E())
    assert close(ens.elastic_energy()[1], o.poly_E())
    # the same two sweeps on the CPU, from the same production streams
    o2 = O.OracleSim(spec0)
    o2.use_production_streams(41, 1)
    mv = O.make_moves(400_000, 16.5)
    o2.mc_sim(mv, 2, 0)
    assert [int(x) for x in ens.moves["num_success"][1]] == [m.num_success for m in mv]
    assert np.array_equal(ens.states[1], o2.states)
    assert np.allclose(ens.r[1], o2.r, rtol=0, atol=1e-7)
    ens.close()


def test_replay_matches_oracle_at_c2_size(cuda_backend):
    """A replayed mc_sim at N = 10,000 reproduces the oracle's accept/reject
    counts and final configuration (the oracle needs ~0.2 s per 1,000 attempts here)."""
    ens, g, binders, Rc = _ensemble(2, 10_000, nb=1, seed=4)
    spec = _spec_of(ens, g, binders, 1)
    o = O.OracleSim(spec)
    mv = O.make_moves(10_000, 16.5)
    o.srand(9)
    o.mc_sim(mv, 5, 77)
    ens.engine.srand(9)
    ens.mc_sim(5, 1.0, 77, rng="replay", sync_host=True)
    assert [int(x) for x in ens.moves["num_success"][1]] == [m.num_success for m in mv]
    assert np.allclose(ens.r[1], o.r, rtol=0, atol=1e-7)
    assert np.array_equal(ens.states[1], o.states)
    assert np.array_equal([float(x) for x in ens.moves["amp_move"][1]], [m.amp_move for m in mv])
    ens.close()


def _check_against_production_oracle(ens, g, binders, start, reps, sweeps, seed, offset=0):
    """replicas `reps` of a production-mode run against the oracle drawing from the same streams"""
    for rep in reps:
        spec = dict(start["spec"](rep))
        o = O.OracleSim(spec)
        o.use_production_streams(seed, offset + rep)
        mv = O.make_moves(ens.N, 16.5)
        o.mc_sim(mv, sweeps, 0)
        assert [int(x) for x in ens.moves["num_success"][rep]] == [m.num_success for m in mv], rep
        assert [int(x) for x in ens.moves["amp_bead"][rep]] == [m.amp_bead for m in mv]
        assert [float(x) for x in ens.moves["amp_move"][rep]] == [m.amp_move for m in mv]
        assert np.array_equal(ens.states[rep], o.states)
        assert np.allclose(ens.r[rep], o.r, rtol=0, atol=1e-7)
        assert np.allclose(ens.t3[rep], o.t3, rtol=0, atol=1e-7)


def test_production_run_matches_oracle_at_c2_size(cuda_backend):
    """The benchmarked instantiation (Philox streams, batches of 32 prepared attempts, 7 replicas per block,
    default table) at N = 10,000 for 6 sweeps: same accept / reject counts, controller state, binding states
    and positions as the oracle run sequentially on the same streams."""
    R = 16
    ens, g, binders, Rc = _ensemble(R, 10_000, nb=1, seed=6)
    assert ens.engine.set_replicas_per_block(7) == 7
    r0, t30, t20, st0 = ens.r.copy(), ens.t3.copy(), ens.t2.copy(), ens.states.copy()
    start = dict(spec=lambda rep: dict(N=10_000, nb=1, r=r0[rep], t3=t30[rep], t2=t20[rep], states=st0[rep],
                                       mods=ens.chemical_mods[rep], bead_length=np.full(9_999, 16.5), lp=53.0,
                                       bead_rad=5.0, binders=binders, max_binders=-1, field=dict(g, chi=1.0)))
    ens.mc_sim(6, 1.0, 31337, sync_host=True)
    _check_against_production_oracle(ens, g, binders, start, (0, 6, 7, 15), 6, 31337)
    ens.close()


def test_bench_ensemble_replays_against_oracle(cuda_backend):
    """The bench's own ensemble (bench.make_inputs: 1,024 replicas x 10,000 beads, marks from the reference's
    H3K9me3 track, default launch shape: 7 replicas per block, 147 blocks) for 5 sweeps; three of its replicas
    replayed on the CPU."""
    import bench
    from chromo_b200.ensemble import ReplicaEnsemble, default_moves
    R, N = 1024, 10_000
    r, t3, t2, states, mods, grid = bench.make_inputs(R, N, 1234, pinned=False)
    binders = [dict(bench.HP1)]
    ens = ReplicaEnsemble(r.copy(), t3.copy(), t2.copy(), states.copy(), mods, binders=binders,
                          bond_params=bench.bond_params(N), grid=grid, bead_vol=(4 / 3) * math.pi * 125.0, chi=1.0,
                          mu=[-1.2], moves=default_moves(R, N, 16.5))
    assert ens.engine.set_replicas_per_block(0) == 7
    start = dict(spec=lambda rep: dict(N=N, nb=1, r=r[rep], t3=t3[rep], t2=t2[rep], states=states[rep],
                                       mods=mods[rep], bead_length=np.full(N - 1, 16.5), lp=53.0, bead_rad=5.0,
                                       binders=binders, max_binders=-1, field=dict(grid, chi=1.0)))
    ens.mc_sim(5, 1.0, 2024, sync_host=False)
    ens.sync()
    ens.pull()
    _check_against_production_oracle(ens, grid, binders, start, (0, 511, 1023), 5, 2024)
    ens.close()


def test_replica_exchange_single_rank(cuda_backend):
    """C5 on one rank, decided on the device: four ladders of 8 rungs, labels move, configurations do not; the
    kernel's decisions equal the host statement of the rule; mean Phi decreases with chi along the ladders."""
    from chromo_b200 import parallel as par
    R, L = 32, 8
    ens, g, binders, Rc = _ensemble(R, 2_000, nb=1, seed=5)
    ladder = np.tile(np.geomspace(0.25, 4.0, L), R // L)
    ex = par.ReplicaExchange(ens, ladder, seed=3, ladder_len=L)
    chi = ladder.copy()
    phi_by_rung = np.zeros(L)
    for rnd in range(12):
        ens.mc_sim(20, 1.0, 100 + rnd, sync_host=False)
        ex.step()
        rung, chi_local, tried, acc = ex.state()
        phi = ex.phi_all.cpu().numpy()
        chi = par.swap_decisions(chi, phi, rnd, 3, L)
        assert np.array_equal(chi_local, chi)
        if rnd >= 6:
            for l0 in range(0, R, L):
                phi_by_rung += phi[rung[l0:l0 + L]]
    assert acc > 0 and tried == 6 * 4 * 4 + 6 * 4 * 3
    for l0 in range(0, R, L):
        assert np.array_equal(np.sort(chi[l0:l0 + L]), ladder[:L])
    # a larger chi penalises dense voxels: Phi = sum (V/v) phi^2 falls along the ladder
    assert phi_by_rung[0] > phi_by_rung[-1]
    _invariants(ens, g, Rc)
    ens.close()
