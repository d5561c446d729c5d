"""Parity of the CUDA path with the reference, through the C ABI.

Every test runs twice: on the lockstep CPU emulation of the kernels ("emu",
the not-gpu suite: checks the kernel logic in this container) and on the real
sm_100a library ("cuda", marked gpu).  Golden vectors come from the reference's
own Cython build (tests/golden/make_golden.py); the oracle is
oracle/chromo_oracle.c.

Bars (north_star): bin indices / touched-bin sets bit-exact; energies and dE
within 1e-9 relative in fp64; with replayed RNG draws identical accept/reject
sequences.  Trial coordinates are bit-exact under emulation (same libm) and
within 1e-9 nm on the GPU (CUDA's sin/cos/acos/log10 differ from glibc's in the
last ulp)."""
import ctypes as C

import numpy as np
import pytest

from common import close, close_dE, huge_scale, load_golden, split
from gpu_common import engine_from_spec, moves_array

STATIC = ["static_c1", "static_c2", "static_c3", "static_c4", "static_tw", "static_tw2", "static_av", "static_av2", "static_dc"]
MOVES = ["moves_c1", "moves_c2", "moves_c3", "moves_c4", "moves_tw", "moves_tw2", "moves_av", "moves_av2", "moves_ff", "moves_ff2", "moves_dc", "moves_dc2", "moves_dc3"]
MCSIM = ["mcsim_c1", "mcsim_c2", "mcsim_c3", "mcsim_tw", "mcsim_av", "mcsim_ff", "mcsim_dc", "mcsim_dc3", "mcsim_order"]
REPLAY, PHILOX = 1, 0


@pytest.mark.parametrize("name", STATIC)
def test_full_recompute_and_total_energies(backend, name):
    """A8: update_all_densities, compute_E (field) and SSWLC.compute_E."""
    spec, g = load_golden(name)
    e = engine_from_spec(spec, R=3)
    d = e.density()
    for rep in range(3):
        assert np.array_equal(d[rep] != 0, g["density"] != 0)  # occupied-bin set, bit-exact
        assert np.allclose(d[rep], g["density"], rtol=1e-12, atol=0)
    E, sq, dbl, ns = e.field_energy()
    Ep = e.elastic_energy()
    for rep in range(3):
        assert close(E[rep], float(g["E_field"]))
        assert close(Ep[rep], float(g["E_poly"]))
    vols = g["access_vols"] if "access_vols" in g else float(g["vol_bin"])  # per-voxel volumes of the *_av cases
    assert np.allclose((d[..., 0] * vols).sum(axis=1), spec["N"], rtol=1e-12)  # mass conservation
    e.close()


@pytest.mark.parametrize("name", MOVES)
def test_single_move_chain(backend, name):
    """A1-A7, A9-A11: a chain of mc_steps replaying the reference's RNG draws:
    proposal indices, trial rows, elastic dE, field dE, touched bins,
    density_trial rows; accept/reject forced to the golden decision."""
    exact = backend == "emu"
    spec, g = load_golden(name)
    e = engine_from_spec(spec, R=2)
    seed = int(g["seed"])
    e.srand(seed)
    e.numpy_seed(seed)
    inds_l = split(g["inds"], g["n"])
    touched_l = split(g["touched"], g["n_touched"])
    dtrial_l = split(g["dtrial"], g["n_touched"])
    rows_l = split(g["trial_rows"], g["n"])
    vol_bin = spec["field"]["x_width"] ** 3 / spec["field"]["nx"] ** 3
    bead_vol = (4 / 3) * np.pi * spec["bead_rad"] ** 3
    for it in range(len(g["move"])):
        m = int(g["move"][it])
        dens = e.density(1, 1)[0] if (m != 3 and spec["field"]["vf_limit"] < 0.5) else None
        out = e.mc_step(1, m, float(g["amp_move"][it]), int(g["amp_bead"][it]), 1.0, REPLAY, 0,
                        int(g["accept"][it]))
        assert np.array_equal(out["inds"], inds_l[it]), (it, m)
        if exact:
            assert np.array_equal(out["rows"], rows_l[it]), (it, m)
        else:
            assert np.allclose(out["rows"], rows_l[it], rtol=1e-12, atol=1e-9), (it, m)
        assert close(out["dE_poly"], float(g["dE_poly"][it]), 1e-9, 1e-9), (it, m)
        if m != 3:
            order = np.argsort(out["touched"])
            assert np.array_equal(out["touched"][order], touched_l[it]), (it, m)  # bit-exact set
            assert np.allclose(out["dtrial"][order], dtrial_l[it], rtol=1e-9, atol=1e-9 / vol_bin), (it, m)
            sc = 0.0
            if dens is not None:
                sc = huge_scale(dens, _scatter(dens.shape, touched_l[it], dtrial_l[it]), touched_l[it],
                                bead_vol, spec["field"]["vf_limit"])
            assert close_dE(out["dE_field"], float(g["dE_field"][it]), sc), (it, m)
        assert out["accepted"] == bool(g["accept"][it])
    r, t3, t2, st = e.download()
    tol, rtol = (0, 0) if exact else (1e-8, 1e-12)
    assert np.allclose(r[1], g["final_r"], rtol=rtol, atol=tol)
    assert np.allclose(t3[1], g["final_t3"], rtol=rtol, atol=tol)
    assert np.allclose(t2[1], g["final_t2"], rtol=rtol, atol=tol)
    assert np.array_equal(st[1], g["final_states"])
    assert np.allclose(e.density(1, 1)[0], g["final_density"], rtol=1e-9, atol=1e-9 / vol_bin)
    # replica 0 was never stepped
    assert np.array_equal(r[0], spec["r"])
    e.close()


def _scatter(shape, touched, rows):
    a = np.zeros(shape)
    a[touched] = rows
    return a


@pytest.mark.parametrize("name", MCSIM)
def test_mc_sim_replay(backend, name):
    """A12 + the whole loop: a full mc_sim under the reference's RNG streams
    reproduces the reference's accept/reject sequence, controller state and
    final configuration."""
    exact = backend == "emu"
    spec, g = load_golden(name)
    R = 2
    e = engine_from_spec(spec, R=R)
    mv = moves_array(spec, R, tuple(int(x) for x in g["per_cycle"]))
    e.srand(int(g["srand_seed"]))
    if "order" in g:  # the reference's controller list was not in all_moves' order
        e.set_move_order(g["order"])
    e.mc_sim(int(g["steps"]), mv, float(g["mu_adjust"]), 0, REPLAY, numpy_seeds=int(g["np_seed"]))
    r, t3, t2, st = e.download()
    for rep in range(R):
        assert list(mv["num_attempt"][rep]) == list(g["num_attempt"])
        assert list(mv["num_success"][rep]) == list(g["num_success"])  # same accept/reject sequence
        assert list(mv["amp_bead"][rep]) == list(g["amp_bead"])
        assert np.array_equal(mv["amp_move"][rep], g["amp_move"])
        assert np.array_equal(mv["acceptance_rate"][rep], g["acceptance_rate"])
        tol = 0 if exact else 1e-7
        assert np.allclose(r[rep], g["final_r"], rtol=0, atol=tol)
        assert np.allclose(t3[rep], g["final_t3"], rtol=0, atol=tol)
        assert np.array_equal(st[rep], g["final_states"])
    assert e.last_attempts() == R * int(np.sum(g["num_attempt"]))
    vol_bin = spec["field"]["x_width"] ** 3 / spec["field"]["nx"] ** 3
    assert np.allclose(e.density()[0], g["final_density"], rtol=1e-9, atol=1e-9 / vol_bin)
    E, _, _, _ = e.field_energy()
    assert close(E[0], float(g["E_field"]), 1e-9 if exact else 1e-7)
    assert close(e.elastic_energy()[R - 1], float(g["E_poly"]), 1e-9 if exact else 1e-7)
    e.close()


def test_large_moves_take_partition_passes(backend, oracle_mod):
    """Moves whose touched-voxel set overflows the shared-memory table are
    evaluated in hash-partition passes; results must not change."""
    O = oracle_mod
    spec = O.make_spec(N=300, nb=2, seed=5, grid=24, cross_talk=-0.7)
    e = engine_from_spec(spec, R=1)
    assert e.set_table_capacity(128) == 128
    o = O.OracleSim(spec)
    e.srand(3), o.srand(3), e.numpy_seed(3), o.np_seed(3)
    mvs = O.make_moves(spec["N"], 16.5)
    rng = np.random.default_rng(0)
    maxp = 0
    for it in range(40):
        m = int(rng.integers(0, 5))
        amp_bead, amp_move = int(rng.integers(40, 150)), 0.3 * (1 + m)
        inds = o.propose(m, amp_move, amp_bead)
        dEp = o.poly_dE(m, inds)
        dEf, touched = (0.0, np.zeros(0, dtype=np.int64)) if m == 3 else o.field_dE(inds, m == 4)
        with np.errstate(over="ignore"):
            acc = int(rng.uniform() < np.exp(-(dEp + dEf)))
        out = e.mc_step(0, m, amp_move, amp_bead, 1.0, REPLAY, 0, acc)
        maxp = max(maxp, out["passes"])
        assert np.array_equal(out["inds"], inds)
        assert close(out["dE_poly"], dEp, 1e-9, 1e-9)
        if m != 3:
            assert np.array_equal(np.sort(out["touched"]), np.sort(touched))
            a = out["dtrial"][np.argsort(out["touched"])]
            b = o.density_trial[np.sort(touched)]
            assert np.allclose(a, b, rtol=1e-9, atol=1e-9 / o.s.vol_bin)
            # voxels above vf_limit contribute +-1e99*phi terms that cancel to ~1e-16 of themselves
            sc = huge_scale(o.density, o.density_trial, np.sort(touched), o.s.bead_vol, spec["field"]["vf_limit"])
            assert close_dE(out["dE_field"], dEf, sc)
        ip = inds.ctypes.data_as(O._pl)
        if acc:
            O.lib().oc_accept(C.byref(o.s), C.byref(mvs[m]), m, ip, len(inds))
            if m != 3:
                o.commit_field()
        else:
            O.lib().oc_reject(C.byref(o.s), C.byref(mvs[m]), m, ip, len(inds))
    assert maxp >= 4
    r, t3, t2, st = e.download()
    assert np.allclose(r[0], o.r, rtol=0, atol=1e-8) and np.array_equal(st[0], o.states)
    assert np.allclose(e.density()[0], o.density, rtol=1e-9, atol=1e-9 / o.s.vol_bin)
    e.close()


def test_binning_edge_cases(backend, oracle_mod):
    """Beads exactly on voxel faces, on the box faces, several box widths away
    and a hair below zero (Python-modulo result == W): bin sets bit-exact."""
    O = oracle_mod
    spec = O.make_spec(N=64, nb=1, seed=9, confine="", grid=6, random_states=True)
    W = spec["field"]["x_width"]
    d = W / 6
    pts = [0.0, -0.0, d, -d, d / 2, -d / 2, W / 2, -W / 2, W, -W, 3 * W + d / 2, -5 * W - d / 2,
           np.nextafter(-W / 2, -np.inf), np.nextafter(-W / 2, np.inf), -W / 2 - 1e-17, 1e-300, -1e-300,
           np.nextafter(W / 2, np.inf), np.nextafter(d / 2, 0), np.nextafter(d / 2, W)]
    rng = np.random.default_rng(1)
    r = rng.choice(pts, size=(64, 3))
    spec["r"] = r
    e = engine_from_spec(spec, R=1)
    o = O.OracleSim(spec)
    dd = e.density()[0]
    assert np.array_equal(dd != 0, o.density != 0)
    assert np.allclose(dd, o.density, rtol=1e-12, atol=0)
    E, _, _, _ = e.field_energy()
    assert close(E[0], o.field_E())
    e.close()


def test_philox_is_deterministic_and_consistent(backend, oracle_mod):
    """Production RNG: same seed -> identical trajectory; different replicas
    decorrelate; the incrementally updated density equals a full recompute;
    mass is conserved; states stay within [0, sites]."""
    O = oracle_mod
    spec = O.make_spec(N=150, nb=2, seed=12, cross_talk=-0.5, random_states=False)
    R = 3
    runs = []
    for _ in range(2):
        e = engine_from_spec(spec, R=R)
        mv = moves_array(spec, R)
        e.mc_sim(2 if backend == "emu" else 4, mv, 1.0, 77, PHILOX)
        r, t3, t2, st = e.download()
        dens = e.density()
        runs.append((r, st, dens, mv.copy()))
        if _ == 0:
            e.field_recompute(clamp=False)
            assert np.allclose(e.density(), dens, rtol=1e-9, atol=1e-12 * dens.max())
            o = O.OracleSim(spec)
            assert np.allclose(dens[..., 0].sum(axis=1) * o.s.vol_bin, spec["N"], rtol=1e-9)
            assert st.min() >= 0 and st.max() <= 2
            assert np.all(np.linalg.norm(r, axis=2) <= spec["field"]["confine_length"] + 1e-9)
            n3 = np.linalg.norm(t3, axis=2)
            assert np.allclose(n3, 1.0, atol=1e-9)
        e.close()
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    # the full recompute (shared-memory privatised, fixed point) and the MC kernel's own updates are both
    # independent of the order in which lanes arrive: the densities of two runs are identical to the last bit
    assert np.array_equal(runs[0][2], runs[1][2])
    assert not np.array_equal(runs[0][0][0], runs[0][0][1])
    acc = runs[0][3]["num_success"].sum() / runs[0][3]["num_attempt"].sum()
    assert 0.2 < acc < 0.95


@pytest.mark.parametrize("case", ["mixed", "dense", "no_field", "twist"])
def test_two_warps_per_replica_match_one(backend, oracle_mod, case):
    """The production kernel with two warps per replica (stage 1 of attempt j+1
    overlapping stage 2 of attempt j, redone when an accepted attempt changed
    rows it had read) gives the sequential result bit for bit.  "dense": a
    short chain with wide windows, so nearly every look-ahead is stale;
    "no_field": only confinement (NullField path)."""
    O = oracle_mod
    if case == "mixed":
        spec = O.make_spec(N=400, nb=2, seed=21, cross_talk=-0.5, random_states=True)
        sweeps, per_cycle = 6, (30, 1, 60, 60, 10)
    elif case == "dense":
        spec = O.make_spec(N=40, nb=1, seed=22, random_states=True)
        sweeps, per_cycle = 8, (30, 5, 60, 60, 30)
    elif case == "twist":  # the SSTWLC builds of the kernels
        spec = dict(O.make_spec(N=150, nb=1, seed=24, random_states=True), lt=80.0)
        sweeps, per_cycle = 4, (30, 1, 60, 60, 10)
    else:
        spec = O.make_spec(N=120, nb=1, seed=23, random_states=False)
        spec["field"] = dict(spec["field"], nx=0, ny=0, nz=0)
        sweeps, per_cycle = 6, (30, 1, 60, 60, 10)
    R = 3
    out = []
    dens0 = None
    for warps in (1, 2):
        e = engine_from_spec(spec, R=R)
        assert e.set_warps_per_replica(warps) == warps
        # one replica per block vs. all three sharing a block (move types entered together)
        assert e.set_replicas_per_block(1 if warps == 1 else R) == (1 if warps == 1 else R)
        if case != "no_field":  # same starting density, bit for bit (the full recompute is atomics-ordered)
            if dens0 is None:
                dens0 = e.density()
            e.upload_density(dens0)
        mv = moves_array(spec, R, per_cycle)
        if case == "dense":
            mv["amp_bead"][:, 3] = 26  # tangent rotation: bead sets on both sides of CB_KSEL = 24 (prepared /
            mv["bead_amp_hi"][:, 3] = 36  # sequential path)
            mv["amp_bead"][:, 4] = 5
            mv["bead_amp_hi"][:, 4] = 5
        e.mc_sim(sweeps, mv, 1.0, 4242, PHILOX)
        r, t3, t2, st = e.download()
        out.append((r, t3, t2, st, e.density() if case != "no_field" else None, mv.copy(), e.last_attempts()))
        e.close()
    a, b = out
    for x, y in zip(a[:4], b[:4]):
        assert np.array_equal(x, y)
    if case != "no_field":
        assert np.array_equal(a[4], b[4])
    for f in ("num_attempt", "num_success", "amp_bead", "amp_move", "acceptance_rate"):
        assert np.array_equal(a[5][f], b[5][f]), f
    assert a[6] == b[6] == R * sweeps * sum(per_cycle)
    assert a[5]["num_success"].sum() > 0


def test_error_behaviour(backend, oracle_mod):
    """Argument checks mirror the reference's exceptions."""
    from chromo_b200 import _lib
    from chromo_b200.engine import Engine
    O = oracle_mod
    spec = O.make_spec(N=50, nb=1, seed=2)
    e = engine_from_spec(spec, R=1)
    mv = moves_array(spec, 1)
    mv["amp_bead"][0, 0] = 51  # bead_selection.pyx:85-88: window larger than the chain
    with pytest.raises(_lib.ChromoError, match="window size"):
        e.mc_sim(1, mv, 1.0, 0, PHILOX)
    with pytest.raises(_lib.ChromoError):
        e.mc_step(5, 0, 0.1, 3)
    with pytest.raises(ValueError, match="Confinement type"):
        Engine(1, 10, 1, grid=dict(spec["field"], confine_type="Ellipsoidal"), bead_vol=1.0)
    bad = spec["states"].copy()
    bad[3, 0] = 3  # above sites_per_bead = 2: would index past the binding free-energy table
    with pytest.raises(_lib.ChromoError, match="must lie in"):
        e.upload(spec["r"][None], spec["t3"][None], spec["t2"][None], bad[None], spec["mods"][None])
    e.upload(spec["r"][None], spec["t3"][None], spec["t2"][None], spec["states"][None], spec["mods"][None])
    e2 = Engine(1, 10, 1, grid=None, bead_vol=1.0)
    with pytest.raises(_lib.ChromoError, match="set_binders"):
        e2.mc_sim(1, None, 1.0, 0, PHILOX)
    e.close(), e2.close()


def test_host_array_path_matches_resident_path(backend, oracle_mod):
    """chromo_mc_sim_host (host arrays in and out, pipelined over replica chunks) reproduces
    upload + mc_sim + download bit for bit, for any chunking: replicas have their own streams."""
    from chromo_b200.ensemble import ReplicaEnsemble, default_moves
    O = oracle_mod
    R, N = (5 if backend == "emu" else 9), 60
    specs = [O.make_spec(N=N, nb=1, seed=50 + i) for i in range(R)]
    st = lambda k: np.stack([s[k] for s in specs])
    kw = dict(binders=specs[0]["binders"], bond_params=O.bond_params(specs[0]["bead_length"], 53.0),
              grid=specs[0]["field"], bead_vol=(4 / 3) * np.pi * 5.0 ** 3, chi=1.0)
    out = {}
    for n_chunks in (-1, 0, 3):           # unpipelined; automatic; 3 ragged chunks (5 blocks of 2 replicas)
        ens = ReplicaEnsemble(st("r"), st("t3"), st("t2"), st("states"), st("mods"),
                              moves=default_moves(R, N, 16.5), **kw)
        ens.engine.set_replicas_per_block(2)
        for k in range(2):
            ens.mc_sim(1 if backend == "emu" else 2, 1.0, 77 + k, n_chunks=n_chunks)
        out[n_chunks] = (ens.r.copy(), ens.t3.copy(), ens.t2.copy(), ens.states.copy(), ens.moves.copy(),
                         ens.density().copy())
        assert ens.chemical_mods.tobytes() == st("mods").tobytes()
        if n_chunks == 3:  # argument checking of the host path: dtype / layout / writability
            with pytest.raises(ValueError):
                ens.engine.mc_sim_host(1, ens.r.astype(np.float32), ens.t3, ens.t2, ens.states, ens.chemical_mods)
            with pytest.raises(ValueError):
                ens.engine.mc_sim_host(1, ens.r[:, ::2], ens.t3, ens.t2, ens.states, ens.chemical_mods)
            ro = ens.states.copy()
            ro.setflags(write=False)
            with pytest.raises(ValueError):
                ens.engine.mc_sim_host(1, ens.r, ens.t3, ens.t2, ro, ens.chemical_mods)
        ens.close()
    for n_chunks in (0, 3):
        for a, b in zip(out[-1][:5], out[n_chunks][:5]):
            assert a.tobytes() == b.tobytes()
        # full recompute and incremental updates are order-independent: identical densities for any chunking
        assert np.array_equal(out[-1][5], out[n_chunks][5])
    assert not np.array_equal(out[-1][0], st("r"))
