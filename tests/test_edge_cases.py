"""Edge cases of the path against the oracle, on both backends: tiny chains,
three and four binders, cubical confinement, max_binders, positive chemical
potentials with annealing, moves switched off, NoControl, huge bead windows
(tangent rotation of > 32 beads, segments that overflow the table)."""
import numpy as np
import pytest

import oracle as O
from common import close
from gpu_common import engine_from_spec, moves_array

REPLAY = 1


def _replay_both(spec, steps, per_cycle=(30, 1, 60, 60, 10), controller=1, move_on=(1, 1, 1, 1, 1),
                 mu_adjust=1.0, tweak=None, srand=5, npseed=9, backend="emu"):
    e = engine_from_spec(spec, R=1)
    mv = moves_array(spec, 1, per_cycle, controller, move_on)
    omv = O.make_moves(spec["N"], float(np.min(spec["bead_length"])), per_cycle=per_cycle,
                       controller=controller, move_on=move_on)
    if tweak:
        tweak(mv, omv)
    o = O.OracleSim(spec, mu_adjust_factor=mu_adjust)
    o.srand(srand)
    o.mc_sim(omv, steps, npseed)
    e.srand(srand)
    e.mc_sim(steps, mv, mu_adjust, 0, REPLAY, numpy_seeds=npseed)
    r, t3, t2, st = e.download()
    assert [int(x) for x in mv["num_attempt"][0]] == [m.num_attempt for m in omv]
    assert [int(x) for x in mv["num_success"][0]] == [m.num_success for m in omv]
    assert [int(x) for x in mv["amp_bead"][0]] == [m.amp_bead for m in omv]
    tol = 0 if backend == "emu" else 1e-7
    assert np.allclose(r[0], o.r, rtol=0, atol=tol) and np.allclose(t3[0], o.t3, rtol=0, atol=tol)
    assert np.allclose(t2[0], o.t2, rtol=0, atol=tol)
    assert np.array_equal(st[0], o.states)
    if spec["field"] is not None:
        assert np.allclose(e.density()[0], o.density, rtol=1e-9, atol=1e-9 / o.s.vol_bin)
        assert close(e.field_energy()[0][0], o.field_E(), 1e-9 if backend == "emu" else 1e-7)
    assert close(e.elastic_energy()[0], o.poly_E(), 1e-9 if backend == "emu" else 1e-7)
    e.close()
    return mv, omv


@pytest.mark.parametrize("N", [2, 3, 5])
def test_tiny_chains(backend, N):
    """Shortest chains: every crank-shaft / pivot axis special case
    (move_funcs.pyx:200-226, 260-280, 384-398) is hit."""
    spec = O.make_spec(N=N, nb=1, seed=N, grid=4, confine="")  # the sphere for N=2 is smaller than one bond
    # get_amplitude_bounds gives end_pivot a window of min(50, N/4) < 1 for N < 4, for which the
    # reference's capped_exponential never terminates (bead_selection.pyx:62-65): pivot off there
    _replay_both(spec, 25, per_cycle=(6, 4, 6, 6, 3), move_on=(1, int(N >= 4), 1, 1, 1), backend=backend)


def test_three_and_four_binders(backend):
    b3 = [dict(O.HP1, name="HP1", cross_talk={"PRC1": -0.5}), dict(O.PRC1, cross_talk={"HP1": 0.0}),
          dict(O.HP1, name="null_reader", sites_per_bead=1, interaction_energy=-1.0, cross_talk={})]
    for b in b3:
        b["chemical_potential"] = -0.8
    spec = O.make_spec(N=90, nb=3, seed=31, binders=b3)
    _replay_both(spec, 6, backend=backend)
    b4 = b3 + [dict(O.PRC1, name="X4", sites_per_bead=3, bind_energy_mod=-0.3, chemical_potential=-0.4,
                    cross_talk={"HP1": 0.7})]
    spec = O.make_spec(N=70, nb=4, seed=32, binders=b4)
    _replay_both(spec, 5, backend=backend)


def test_cubical_confinement_and_max_binders(backend):
    spec = O.make_spec(N=80, nb=2, seed=33, cross_talk=-0.3, max_binders=2)
    f = spec["field"]
    f["confine_type"], f["confine_length"] = "Cubical", 2 * f["confine_length"]
    _replay_both(spec, 8, backend=backend)


def test_positive_mu_and_annealing(backend):
    spec = O.make_spec(N=100, nb=1, seed=34, mu=0.6)
    _replay_both(spec, 6, mu_adjust=1.7, backend=backend)
    spec = O.make_spec(N=100, nb=1, seed=35, mu=-0.6)
    _replay_both(spec, 6, mu_adjust=3.0, backend=backend)


def test_moves_off_and_no_control(backend):
    spec = O.make_spec(N=120, nb=1, seed=36)
    mv, omv = _replay_both(spec, 5, move_on=(1, 0, 1, 0, 1), backend=backend)
    assert mv["num_attempt"][0, 1] == 0 and mv["num_attempt"][0, 3] == 0
    mv, omv = _replay_both(spec, 5, controller=0, backend=backend)
    assert np.array_equal(mv["amp_move"][0], [m.amp_move for m in O.make_moves(120, 16.5)])


def test_null_field(backend):
    """NullField: elastic + binding energies only (tutorial 1)."""
    spec = O.make_spec(N=150, nb=1, seed=37)
    spec["field"] = None
    _replay_both(spec, 8, backend=backend)


def test_wide_windows(backend):
    """Bead windows as wide as the chain: tangent rotations of > 32 beads (the
    bitmap / HBM-scratch path) and segments of > 100 beads (partition passes)."""
    spec = O.make_spec(N=160, nb=1, seed=38, grid=14)

    def tweak(mv, omv):
        for i, (ab, hi) in enumerate(((150, 160), (75, 80), (140, 160), (120, 160), (40, 40))):
            mv["amp_bead"][0, i] = ab
            mv["bead_amp_hi"][0, i] = hi
            mv["bead_amp_lo"][0, i] = min(mv["bead_amp_lo"][0, i], ab)
            omv[i].amp_bead, omv[i].bead_amp_hi = ab, hi
            omv[i].bead_amp_lo = min(omv[i].bead_amp_lo, ab)
    e_cap = {}
    _replay_both(spec, 4, per_cycle=(8, 4, 8, 12, 6), tweak=tweak, backend=backend)


@pytest.mark.parametrize("N", [2, 5, 60])
def test_twisted_chains(backend, N):
    """SSTWLC (twist term in every bond energy, polymers.pyx:1889-2319) on the shortest chains -- every
    end-of-chain special case with the trial t2 in play -- and on a chain with wide windows; replayed
    against the oracle."""
    spec = dict(O.make_spec(N=N, nb=1, seed=40 + N, grid=4, confine=""), lt=70.0)

    def wide(mv, omv):
        if N >= 60:
            for i, a in ((0, 40), (2, 30), (3, 25)):
                mv["amp_bead"][:, i] = a
                omv[i].amp_bead = a
    _replay_both(spec, 12, per_cycle=(8, 4, 8, 10, 3), move_on=(1, int(N >= 4), 1, 1, 1), tweak=wide, backend=backend)


def test_metropolis_flips_exactly_at_the_accept_boundary(backend):
    """north_star: accept / reject sequences must match "except where |dE - ln u| falls within" 1e-9.  A binder
    move's dE is linear in mu_adjust_factor (polymers.pyx:1493-1517), its proposal and its uniform u are not touched
    by it: solve for the factor a* that puts dE on the boundary dE = -ln u, then (i) a clear step to either side
    (1e-6 relative, far outside the tolerance) must flip the decision on the kernel exactly as on the oracle, and
    (ii) within the tolerance the two may disagree, but only there -- and a replay that follows the kernel's
    decision stays in step with it."""
    spec = O.make_spec(N=120, nb=1, seed=33, random_states=True)

    def attempt(a, follow=None):
        e = engine_from_spec(spec, R=1)
        o = O.OracleSim(spec, mu_adjust_factor=a)
        e.srand(8), o.srand(8), e.numpy_seed(4), o.np_seed(4)
        out = e.mc_step(0, 4, 1.0, 40, a, REPLAY, 0, -1)
        inds = o.propose(4, 1.0, 40)
        assert np.array_equal(inds, out["inds"])
        dE = o.poly_dE(4, inds) + o.field_dE(inds, True)[0]
        u = O.lib().oc_rand(O.C.byref(o.s.crng)) / 2147483647.0  # `<double>rand() / RAND_MAX`, mc_sim.pyx:171
        assert u == out["u"]
        with np.errstate(over="ignore"):
            o_acc = bool(u < np.exp(-dE))
        k_dE = out["dE_poly"] + out["dE_field"]
        if follow is not None:  # replay convention inside the tolerance: take the kernel's decision
            mv = O.make_moves(spec["N"], 16.5)
            ip = inds.ctypes.data_as(O._pl)
            if out["accepted"]:
                O.lib().oc_accept(O.C.byref(o.s), O.C.byref(mv[4]), 4, ip, len(inds))
                o.commit_field()
            else:
                O.lib().oc_reject(O.C.byref(o.s), O.C.byref(mv[4]), 4, ip, len(inds))
            st = e.download()[3]
            assert np.array_equal(st[0], o.states)
            assert np.allclose(e.density()[0], o.density, rtol=1e-9, atol=1e-9 / o.s.vol_bin)
        e.close()
        return dict(u=u, dE=dE, k_dE=k_dE, o_acc=o_acc, k_acc=out["accepted"])

    a1, a2 = attempt(1.0), attempt(2.0)
    assert a1["u"] == a2["u"] and a1["dE"] != a2["dE"]  # same proposal and uniform, dE moved with the factor
    slope = a2["dE"] - a1["dE"]
    target = -np.log(a1["u"])
    a_star = 1.0 + (target - a1["dE"]) / slope
    on = attempt(a_star, follow=True)
    tol = 1e-9 * max(1.0, abs(target))
    assert abs(on["dE"] - target) <= tol and abs(on["k_dE"] - target) <= tol  # both sit on the boundary
    # (i) clear of the tolerance: kernel == oracle == the side of the boundary
    step = 1e-6 * max(1.0, abs(target)) / abs(slope)
    up, dn = attempt(a_star + step * np.sign(slope)), attempt(a_star - step * np.sign(slope))
    assert abs(up["dE"] - target) > 100 * tol and abs(dn["dE"] - target) > 100 * tol
    assert (up["k_acc"], up["o_acc"]) == (False, False)  # dE above -ln u: rejected
    assert (dn["k_acc"], dn["o_acc"]) == (True, True)    # below: accepted
    # (ii) inside the tolerance a disagreement is allowed, anywhere else it is not
    for da in (0.0, 1e-13 / abs(slope), -1e-13 / abs(slope)):
        w = attempt(a_star + da, follow=True)
        assert abs(w["k_dE"] - w["dE"]) <= tol
        if w["k_acc"] != w["o_acc"]:
            assert abs(w["dE"] - target) <= tol
