"""The CPU oracle (oracle/chromo_oracle.c) against golden vectors produced by
the reference's own Cython build (tests/golden/make_golden.py).

Bit-exact: bins, touched sets, trial rows, proposals, states, final positions
of a replayed mc_sim.  Energies: 1e-9 relative (the only source of difference
is the summation order over the touched-bin set)."""
import ctypes as C

import numpy as np
import pytest

from common import close, close_dE, huge_scale, load_golden, split

STATIC = ["static_c1", "static_c2", "static_c3", "static_c4", "static_tw", "static_tw2", "static_av", "static_av2", "static_dc"]
MOVES = ["moves_c1", "moves_c2", "moves_c3", "moves_c4", "moves_tw", "moves_tw2", "moves_av", "moves_av2", "moves_ff", "moves_ff2", "moves_dc", "moves_dc2", "moves_dc3"]
MCSIM = ["mcsim_c1", "mcsim_c2", "mcsim_c3", "mcsim_tw", "mcsim_av", "mcsim_ff", "mcsim_dc", "mcsim_dc3", "mcsim_order"]


@pytest.mark.parametrize("name", STATIC)
def test_static(oracle_mod, name):
    O = oracle_mod
    spec, g = load_golden(name)
    s = O.OracleSim(spec)
    for k in ("eps_bend", "eps_par", "eps_perp", "gamma", "eta"):
        assert np.array_equal(s.bp[k], g[k]), k
    assert np.array_equal(s.pref, g["field_pref"])
    assert np.array_equal(s.e_intra, g["e_intra"])
    assert np.array_equal(s.xpref, g["xpref"])
    assert s.s.vol_bin == float(g["vol_bin"]) and s.s.bead_vol == float(g["bead_vol"])
    assert np.array_equal(s.density, g["density"])          # bit-exact binning + weights
    assert close(s.field_E(), float(g["E_field"]))
    assert np.array_equal(s.density, g["density_after_E"])
    assert close(s.poly_E(), float(g["E_poly"]))
    # mass conservation (validate_density_calculation.ipynb): sum(rho*V) = N
    assert abs((s.density[:, 0] * s.access_vol).sum() - spec["N"]) < 1e-9 * spec["N"]  # (V = the accessible volume)


@pytest.mark.parametrize("name", MOVES)
def test_move_chain(oracle_mod, name):
    O = oracle_mod
    spec, g = load_golden(name)
    s = O.OracleSim(spec)
    mv = O.make_moves(spec["N"], float(np.min(spec["bead_length"])))
    seed = int(g["seed"])
    s.srand(seed)
    s.np_seed(seed)
    inds_l = split(g["inds"], g["n"])
    touched_l = split(g["touched"], g["n_touched"])
    dtrial_l = split(g["dtrial"], g["n_touched"])
    rows_l = split(g["trial_rows"], g["n"])
    nb = spec["nb"]
    for it in range(len(g["move"])):
        m = int(g["move"][it])
        inds = s.propose(m, float(g["amp_move"][it]), int(g["amp_bead"][it]))
        assert np.array_equal(inds, inds_l[it]), (it, m)
        rows = np.concatenate([s.r_trial[inds], s.t3_trial[inds], s.t2_trial[inds],
                               s.states_trial[inds].astype(float)], axis=1)
        assert np.array_equal(rows, rows_l[it]), (it, m)
        if "lt" in spec:  # twist: the reference's np.dot (BLAS ddot) rounding is unspecified -> 1e-11
            assert close(s.poly_dE(m, inds), float(g["dE_poly"][it]), rtol=1e-11, atol=1e-11), (it, m)
        else:
            assert s.poly_dE(m, inds) == g["dE_poly"][it], (it, m)
        if m != 3:
            dEf, touched = s.field_dE(inds, m == 4)
            tr = np.sort(touched)
            assert np.array_equal(tr, touched_l[it]), (it, m)
            assert np.array_equal(s.density_trial[tr], dtrial_l[it]), (it, m)
            sc = huge_scale(s.density, s.density_trial, tr, s.s.bead_vol, spec["field"]["vf_limit"])
            assert close_dE(dEf, float(g["dE_field"][it]), sc), (it, m, dEf, g["dE_field"][it])
        ip = inds.ctypes.data_as(O._pl)
        if g["accept"][it]:
            O.lib().oc_accept(C.byref(s.s), C.byref(mv[m]), m, ip, len(inds))
            if m != 3:
                s.commit_field()
        else:
            O.lib().oc_reject(C.byref(s.s), C.byref(mv[m]), m, ip, len(inds))
    assert np.array_equal(s.r, g["final_r"])
    assert np.array_equal(s.t3, g["final_t3"])
    assert np.array_equal(s.t2, g["final_t2"])
    assert np.array_equal(s.states, g["final_states"])
    assert np.allclose(s.density, g["final_density"], rtol=1e-12, atol=1e-22)


@pytest.mark.parametrize("name", MCSIM)
def test_mc_sim_replay(oracle_mod, name):
    O = oracle_mod
    spec, g = load_golden(name)
    s = O.OracleSim(spec, mu_adjust_factor=float(g["mu_adjust"]))
    mv = O.make_moves(spec["N"], float(np.min(spec["bead_length"])), per_cycle=[int(x) for x in g["per_cycle"]])
    s.srand(int(g["srand_seed"]))
    s.mc_sim(mv, int(g["steps"]), int(g["np_seed"]), order=g["order"] if "order" in g else None)
    assert np.array_equal(s.r, g["final_r"])
    assert np.array_equal(s.t3, g["final_t3"])
    assert np.array_equal(s.t2, g["final_t2"])
    assert np.array_equal(s.states, g["final_states"])
    assert [m.num_attempt for m in mv] == list(g["num_attempt"])
    assert [m.num_success for m in mv] == list(g["num_success"])
    assert [m.amp_bead for m in mv] == list(g["amp_bead"])
    assert np.array_equal([m.amp_move for m in mv], g["amp_move"])
    assert np.array_equal([m.acceptance_rate for m in mv], g["acceptance_rate"])
    assert np.allclose(s.density, g["final_density"], rtol=1e-12, atol=1e-22)
    assert close(s.field_E(), float(g["E_field"]))
    assert close(s.poly_E(), float(g["E_poly"]))


def test_rng_known_answers(oracle_mod):
    """glibc rand() with seed 1 (a fresh process) and numpy's legacy MT19937."""
    O = oracle_mod
    g = O.GlibcRand()
    O.lib().oc_srand(C.byref(g), 1)
    first = [O.lib().oc_rand(C.byref(g)) for _ in range(5)]
    assert first == [1804289383, 846930886, 1681692777, 1714636915, 1957747793]
    mt = O.MT19937()
    O.lib().oc_mt_seed(C.byref(mt), 5489)
    assert O.lib().oc_mt_next(C.byref(mt)) == 3499211612  # MT19937 reference output
    np.random.seed(12345)
    O.lib().oc_mt_seed(C.byref(mt), 12345)
    a = [int(np.random.randint(0, 3)) for _ in range(500)]
    b = [O.lib().oc_mt_randint(C.byref(mt), 0, 3) for _ in range(500)]
    assert a == b


@pytest.mark.parametrize("name", ["static_av", "static_av2"])
def test_accessible_volumes_match_the_reference(oracle_mod, name):
    """get_accessible_volumes (fields.pyx:714-951) restated in the oracle AND in the product's host code, against
    the reference's own output -- including its floor-divided voxel centres on an even grid."""
    from chromo_b200.fields import accessible_volumes
    spec, g = load_golden(name)
    assert np.array_equal(oracle_mod.accessible_volumes(spec["field"]), g["access_vols"])
    assert np.array_equal(accessible_volumes(spec["field"], 20, 0), g["access_vols"])
    assert (g["access_vols"] != float(g["vol_bin"])).sum() > 20
