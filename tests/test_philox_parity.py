"""The PRODUCTION instantiation of the MC kernel against the oracle.

The kernels that are benchmarked draw from counter-based Philox4x32-10 streams
(one stream per replica and attempt), prepare the state-independent half of up
to 32 attempts at once, patch prepared attempts whose rows an accepted attempt
of the batch changed, and use closed-form sphere points.  None of that runs in
the replayed-reference-stream tests (tests/test_parity.py), so here the oracle
(oracle/chromo_oracle.c, `use_production_streams`) draws from the SAME streams,
strictly sequentially, and the two must agree:

  * whole `mc_sim` runs: identical accept / reject counts and controller state
    per move type, identical binding states, positions within 1e-7 nm on the GPU
    and bit-identical under the CPU emulation (same libm);
  * single attempts through `chromo_mc_step(rng = PHILOX)`: proposal indices and
    touched-voxel sets bit-exact, dE_poly / dE_field within 1e-9;
  * a batch size of 1 gives what a batch size of 32 gives, bit for bit.

Reference lines matched: mc_sim.pyx:139-182 (attempt), move_funcs.pyx:80-99,
229-230, 325-336, 441-451, 503-582, 763-820 (draw order), bead_selection.pyx:19-192.
"""
import ctypes as C

import numpy as np
import pytest

from common import close, close_dE, huge_scale
from gpu_common import engine_from_spec, moves_array

REPLAY, PHILOX = 1, 0


def _case(O, name):
    per_cycle = (30, 1, 60, 60, 10)
    if name == "mixed":      # two binders with cross-talk, random initial states
        spec = O.make_spec(N=300, nb=2, seed=31, cross_talk=-0.5, random_states=True)
    elif name == "dense":    # a short chain with wide windows: nearly every prepared attempt of a batch is stale
        spec = O.make_spec(N=40, nb=1, seed=32, random_states=True)
        per_cycle = (30, 5, 60, 60, 30)
    elif name == "twist":    # SSTWLC builds of the kernels
        spec = dict(O.make_spec(N=300, nb=1, seed=33, random_states=True), lt=80.0)
    elif name == "wide":     # the bench's stationary working point: bead windows at their upper bounds
        spec = O.make_spec(N=700, nb=1, seed=35, random_states=True)
    elif name == "hp1":      # the bench's physics at a small size
        spec = O.make_spec(N=300, nb=1, seed=34, random_states=False)
    else:
        raise KeyError(name)
    return spec, per_cycle


def _dense_moves(mv_arr=None, omv=None):
    """tangent rotations of up to 26 beads, on both sides of the prepared-set limit of 24; binding windows of 5"""
    if mv_arr is not None:
        mv_arr["amp_bead"][:, 3] = 26
        mv_arr["bead_amp_hi"][:, 3] = 36
        mv_arr["amp_bead"][:, 4] = 5
        mv_arr["bead_amp_hi"][:, 4] = 5
    if omv is not None:
        omv[3].amp_bead, omv[3].bead_amp_hi = 26, 36
        omv[4].amp_bead, omv[4].bead_amp_hi = 5, 5


def _wide_moves(mv_arr=None, omv=None):
    """SimpleControl's stationary state on the bench's workload (profiles/stationary_amplitudes.json): segments of
    up to 150 beads (several 32-bead chunks per scatter pass), tangent rotations of ~14"""
    amp = (150, 150, 150, 14, 1)
    for i, a in enumerate(amp):
        if mv_arr is not None:
            mv_arr["amp_bead"][:, i] = a
        if omv is not None:
            omv[i].amp_bead = a


def _compare_run(O, backend, spec, per_cycle, sweeps, R, seed, dense=False, rpb=None, warps=1, batch=None,
                 offset=0, check=None, wide=False):
    exact = backend == "emu"
    e = engine_from_spec(spec, R=R)
    if rpb is not None:
        assert e.set_replicas_per_block(rpb) == rpb
    assert e.set_warps_per_replica(warps) == warps
    if batch is not None:
        e.set_batch_size(batch)
    if offset:
        e.set_replica_offset(offset)
    mv = moves_array(spec, R, per_cycle)
    if dense:
        _dense_moves(mv_arr=mv)
    if wide:
        _wide_moves(mv_arr=mv)
    e.mc_sim(sweeps, mv, 1.0, seed, PHILOX)
    r, t3, t2, st = e.download()
    dens = e.density()
    ctr = e.rng_counters()
    e.close()
    for rep in (range(R) if check is None else check):
        o = O.OracleSim(spec)
        o.use_production_streams(seed, offset + rep)
        omv = O.make_moves(spec["N"], float(np.min(spec["bead_length"])), per_cycle=per_cycle)
        if dense:
            _dense_moves(omv=omv)
        if wide:
            _wide_moves(omv=omv)
        o.mc_sim(omv, sweeps, 0)
        assert [int(x) for x in mv["num_attempt"][rep]] == [m.num_attempt for m in omv]
        assert [int(x) for x in mv["num_success"][rep]] == [m.num_success for m in omv], rep  # same accept sequence
        assert [int(x) for x in mv["amp_bead"][rep]] == [m.amp_bead for m in omv]
        assert [float(x) for x in mv["amp_move"][rep]] == [m.amp_move for m in omv]
        assert [float(x) for x in mv["acceptance_rate"][rep]] == [m.acceptance_rate for m in omv]
        assert np.array_equal(st[rep], o.states)
        tol = 0.0 if exact else 1e-7
        assert np.allclose(r[rep], o.r, rtol=0, atol=tol)
        assert np.allclose(t3[rep], o.t3, rtol=0, atol=tol)
        assert np.allclose(t2[rep], o.t2, rtol=0, atol=tol)
        assert np.allclose(dens[rep], o.density, rtol=1e-9, atol=1e-9 / o.s.vol_bin)
        assert int(ctr[rep]) == o.attempts_made == sweeps * sum(per_cycle)
    return r, t3, t2, st, dens, mv


@pytest.mark.parametrize("name", ["hp1", "mixed", "dense", "twist", "wide"])
def test_production_mc_sim_matches_oracle(backend, oracle_mod, name):
    """(a) whole mc_sim runs in production mode, seven replicas per block (the bench's launch shape)."""
    O = oracle_mod
    spec, per_cycle = _case(O, name)
    emu = backend == "emu"
    sweeps = ({"hp1": 6, "mixed": 4, "dense": 10, "twist": 4, "wide": 3}[name] if emu else
              {"hp1": 12, "mixed": 8, "dense": 20, "twist": 8, "wide": 10}[name])
    R = 9 if emu else 14  # more than one block of 7 replicas: the default launch shape, move-type barrier included
    _compare_run(O, backend, spec, per_cycle, sweeps, R, seed=20241 + len(name), dense=name == "dense", wide=name == "wide",
                 rpb=7, check=range(R) if emu else (0, 6, 7, 13))


def test_production_two_warps_and_offset(backend, oracle_mod):
    """two warps per replica and a non-zero replica offset (a shard of a larger ensemble) draw the streams
    of the global replica indices"""
    O = oracle_mod
    spec, per_cycle = _case(O, "hp1")
    _compare_run(O, backend, spec, per_cycle, 2 if backend == "emu" else 8, 2, seed=99, rpb=2, warps=2, offset=1000)


def test_batch_of_one_is_bit_identical(backend, oracle_mod):
    """(c) the batched preparation (32 attempts ahead, stale ones recomputed) against a batch size of 1, where
    nothing is ever prepared ahead: identical to the last bit on the case where nearly every look-ahead is stale"""
    O = oracle_mod
    spec, per_cycle = _case(O, "dense")
    sweeps = 8 if backend == "emu" else 20
    outs = []
    for batch in (32, 1, 5):
        e = engine_from_spec(spec, R=2)
        if outs:
            e.upload_density(outs[0][4])
        d0 = e.density()
        e.set_batch_size(batch)
        mv = moves_array(spec, 2, per_cycle)
        _dense_moves(mv_arr=mv)
        e.mc_sim(sweeps, mv, 1.0, 777, PHILOX)
        outs.append(e.download() + (d0, e.density(), mv.copy()))
        e.close()
    for other in outs[1:]:
        for a, b in zip(outs[0][:4], other[:4]):
            assert np.array_equal(a, b)
        assert np.array_equal(outs[0][5], other[5])
        for f in ("num_success", "amp_bead", "amp_move", "acceptance_rate"):
            assert np.array_equal(outs[0][6][f], other[6][f])
    assert outs[0][6]["num_success"].sum() > 0


@pytest.mark.parametrize("name", ["hp1", "mixed", "twist"])
def test_production_single_attempts(backend, oracle_mod, name):
    """(b) one attempt at a time through chromo_mc_step(rng = PHILOX): the oracle makes the same attempt from
    the same stream; indices and touched voxels bit-exact, both dE terms within 1e-9, same decision."""
    O = oracle_mod
    spec, _ = _case(O, name)
    exact = backend == "emu"
    seed = 4711
    e = engine_from_spec(spec, R=2)
    o = O.OracleSim(spec)
    o.use_production_streams(seed, 1)
    omv = O.make_moves(spec["N"], 16.5)
    rng = np.random.default_rng(5)
    n_acc = 0
    for it in range(60 if exact else 300):
        m = int(rng.integers(0, 5))
        amp_bead = int(rng.integers(1, 60)) if m != 3 else int(rng.integers(1, 24))
        amp_move = float(rng.uniform(0.05, 1.0)) * (4.0 if m == 2 else 1.0)
        omv[m].amp_move, omv[m].amp_bead = amp_move, amp_bead
        dens_before = o.density.copy()
        rc = o.mc_step(omv[m], m)
        out = e.mc_step(1, m, amp_move, amp_bead, 1.0, PHILOX, seed, -1)
        n = 0 if rc < 0 else int(len(out["inds"]))
        if rc < 0:
            assert len(out["inds"]) == 0
            continue
        inds = o.inds[:n]
        assert np.array_equal(np.sort(out["inds"]), np.sort(inds)), (it, m)
        assert close(out["dE_poly"], o.s.last_dE_poly, 1e-9, 1e-9), (it, m)
        if m != 3:
            touched = np.sort(o.touched[: o.s.n_touched])
            assert np.array_equal(np.sort(out["touched"]), touched), (it, m)  # bit-exact voxel set
            sc = 0.0
            if len(touched):
                dtr = np.zeros_like(dens_before)
                dtr[out["touched"]] = out["dtrial"]
                sc = huge_scale(dens_before, dtr, touched, o.s.bead_vol, spec["field"]["vf_limit"])
            assert close_dE(out["dE_field"], o.s.last_dE_field, sc), (it, m)
        assert out["u"] == o.s.last_u
        assert out["accepted"] == bool(o.s.last_accept), (it, m, out["dE_poly"] + out["dE_field"], out["u"])
        n_acc += int(o.s.last_accept)
    assert n_acc > 5
    r, t3, t2, st = e.download()
    assert np.allclose(r[1], o.r, rtol=0, atol=0 if exact else 1e-8)
    assert np.array_equal(st[1], o.states)
    assert np.array_equal(r[0], spec["r"])  # replica 0 was never stepped
    e.close()
