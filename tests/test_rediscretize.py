"""Coarse-grain / refine pipeline: the CUDA kernels of
csrc/rediscretize.cu through the C ABI and the reference-named host functions,
against (a) golden vectors produced by the reference itself
(tests/golden/rediscretize.npz, made by make_golden_rediscretize.py) and (b) the
numpy restatement in oracle/rediscretize_oracle.py on seeded inputs.

Tolerances: interval means, orientations, majority states -- bit-exact.
Refined paths and confinement -- 1e-12 relative (north_star allows 1e-9): the
reference takes 3-vector norms through BLAS ddot, whose rounding is unspecified.
"""
import numpy as np
import pytest

import rediscretize_oracle as RO
from common import GOLDEN

RTOL = 1e-12


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "rediscretize.npz")


def rel_close(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(1.0, float(np.max(np.abs(b))))
    return a.shape == b.shape and float(np.max(np.abs(a - b))) <= rtol * scale


# ---- the oracle is pinned by the reference's own outputs (CPU) ------------------------------
def test_oracle_matches_reference_golden(gold):
    for i in range(int(gold["cg_n"])):
        N, k, nb = gold[f"cg{i}_shape"]
        r_cg, t3_cg, t2_cg, st, md = RO.cg_arrays(gold[f"cg{i}_r"], gold[f"cg{i}_t3"], gold[f"cg{i}_states"],
                                                  gold[f"cg{i}_mods"], int(k), 1.0)
        assert np.array_equal(r_cg, gold[f"cg{i}_avg"])
        assert np.array_equal(t3_cg, gold[f"cg{i}_t3cg"]) and np.array_equal(t2_cg, gold[f"cg{i}_t2cg"])
        assert np.array_equal(st, gold[f"cg{i}_maj_states"]) and np.array_equal(md, gold[f"cg{i}_maj_mods"])
    for i in range(int(gold["rf_n"])):
        M, n_ref = gold[f"rf{i}_shape"]
        p = RO.refine_path(gold[f"rf{i}_cg"], int(n_ref), float(gold[f"rf{i}_spacing"]), gold[f"rf{i}_xi"])
        assert rel_close(p, gold[f"rf{i}_path"])
        t3, t2 = RO.orient(RO.refine_path(gold[f"rf{i}_t"], int(n_ref), np.pi, gold[f"rf{i}_xo"]))
        assert rel_close(t3, gold[f"rf{i}_t3"]) and rel_close(t2, gold[f"rf{i}_t2"])
    for i in range(int(gold["cf_n"])):
        assert rel_close(RO.confine(gold[f"cf{i}_r"], float(gold[f"cf{i}_rad"])), gold[f"cf{i}_out"])


# ---- the CUDA path against the reference's outputs ----------------------------------------------
def test_coarse_graining_blocks_match_reference(backend, gold):
    import chromo_b200.util.rediscretize as rd
    for i in range(int(gold["cg_n"])):
        N, k, nb = (int(v) for v in gold[f"cg{i}_shape"])
        iv = rd.get_cg_bead_intervals(N, k)
        assert np.array_equal(rd.get_avg_in_intervals(gold[f"cg{i}_r"], iv), gold[f"cg{i}_avg"])
        t3, t2 = rd.get_orientations_in_intervals(gold[f"cg{i}_t3"], iv)
        assert np.array_equal(t3, gold[f"cg{i}_t3cg"]) and np.array_equal(t2, gold[f"cg{i}_t2cg"])
        assert np.array_equal(np.signbit(t2), np.signbit(gold[f"cg{i}_t2cg"]))
        assert np.array_equal(rd.get_majority_state_in_interval(gold[f"cg{i}_states"], iv),
                              gold[f"cg{i}_maj_states"])
        assert np.array_equal(rd.get_majority_state_in_interval(gold[f"cg{i}_mods"], iv), gold[f"cg{i}_maj_mods"])


def _golden_objects(gold):
    from chromo_b200 import binders as bnd, fields, polymers
    N = len(gold["obj_r"])
    df = bnd.make_binder_collection([bnd.get_by_name("HP1")])
    poly = polymers.Chromatin(
        "c", gold["obj_r"].copy(), bead_length=np.ones(N - 1) * 16.5, t3=gold["obj_t3"].copy(),
        t2=gold["obj_t2"].copy(), states=gold["obj_states"].copy(), binder_names=np.array(["HP1"]),
        chemical_mods=gold["obj_mods"].copy(), chemical_mod_names=np.array(["H3K9me3"]))
    R0 = float(gold["obj_R0"])
    udf = fields.UniformDensityField([poly], df, 2.4 * R0, 12, 2.4 * R0, 12, 2.4 * R0, 12,
                                            confine_type="Spherical", confine_length=R0, chi=1.0)
    return poly, df, udf


def _grid_of(u):
    return np.array([u.nx, u.ny, u.nz, u.x_width, u.y_width, u.z_width, u.confine_length])


def test_cg_chromatin_and_field_objects_match_reference(backend, gold):
    import chromo_b200.util.rediscretize as rd
    poly, df, udf = _golden_objects(gold)
    pcg = rd.get_cg_chromatin(poly, int(gold["obj_k"]))
    assert np.array_equal(pcg.r, gold["obj_cg_r"]) and np.array_equal(pcg.t3, gold["obj_cg_t3"])
    assert np.array_equal(pcg.t2, gold["obj_cg_t2"])
    assert np.array_equal(pcg.states, gold["obj_cg_states"])
    assert np.array_equal(pcg.chemical_mods, gold["obj_cg_mods"])
    assert np.array_equal(pcg.bead_length, gold["obj_cg_bead_length"])
    ucg = rd.get_cg_udf(udf.dict_, df, int(gold["obj_k"]), [pcg])
    assert np.array_equal(_grid_of(ucg), gold["obj_cg_grid"])
    assert np.allclose(ucg.density, gold["obj_cg_density"], rtol=1e-12, atol=0)
    # and back: refine_chromatin (geometry; numpy's generator replayed from the same seed)
    np.random.seed(77)
    pref, uref = rd.refine_chromatin(pcg, int(gold["rc_nref"]), 16.5, gold["rc_mods"], ucg)
    assert rel_close(pref.r, gold["rc_r"]) and rel_close(pref.t3, gold["rc_t3"]) and rel_close(pref.t2, gold["rc_t2"])
    assert np.allclose(_grid_of(uref), gold["rc_grid"], rtol=1e-14, atol=0)
    assert np.all(pref.states == 0) and pref.num_beads == int(gold["rc_nref"])


def test_refined_paths_match_reference(backend, gold):
    import chromo_b200.util.rediscretize as rd
    from chromo_b200 import _lib
    for i in range(int(gold["rf_n"])):
        M, n_ref = (int(v) for v in gold[f"rf{i}_shape"])
        L = RO.refine_layout(M, n_ref)
        assert _lib.lib().chromo_refined_num_points(M, n_ref) == L["points"] == len(gold[f"rf{i}_path"])
        assert _lib.lib().chromo_refined_num_draws(M, n_ref) == L["draws"] == len(gold[f"rf{i}_xi"])
        np.random.seed(100 + i)  # the seed the reference was run with
        p = rd.get_refined_path(gold[f"rf{i}_cg"], n_ref, float(gold[f"rf{i}_spacing"]))
        assert rel_close(p, gold[f"rf{i}_path"])
        np.random.seed(200 + i)
        t3, t2 = rd.get_refined_orientations(gold[f"rf{i}_t"], n_ref)
        assert rel_close(t3, gold[f"rf{i}_t3"]) and rel_close(t2, gold[f"rf{i}_t2"])


def test_path_primitives_match_reference(backend, gold):
    """brownian_bridge / gaussian_walk_from_point under the reference's numpy seeds."""
    import chromo_b200.util.rediscretize as rd
    for i in range(int(gold["bb_n"])):
        v = gold[f"bb{i}_in"]
        N, tgt, p0, p1 = int(v[0]), float(v[1]), v[2:5], v[5:8]
        np.random.seed(300 + i)
        assert rel_close(rd.brownian_bridge(N, p0, p1, tgt), gold[f"bb{i}_out"])
        np.random.seed(400 + i)
        assert rel_close(rd.gaussian_walk_from_point(p0, N, np.array([tgt] * N)), gold[f"gw{i}_out"])


def test_confinement_matches_reference(backend, gold):
    import chromo_b200.util.rediscretize as rd
    for i in range(int(gold["cf_n"])):
        r = gold[f"cf{i}_r"].copy()
        out = rd.enforce_spherical_confinement(r, float(gold[f"cf{i}_rad"]))
        assert out is r and rel_close(r, gold[f"cf{i}_out"])


# ---- batched over replicas, against the oracle on seeded inputs ----------------------------------
@pytest.mark.parametrize("R,N,k,nb", [(3, 101, 7, 2), (5, 64, 4, 1), (2, 333, 40, 3)])
def test_ensemble_coarse_graining_matches_oracle(backend, R, N, k, nb):
    import chromo_b200.util.rediscretize as rd
    rng = np.random.default_rng(R * 1000 + N)
    r = np.cumsum(rng.standard_normal((R, N, 3)), axis=1) * 16.5
    t3 = rng.standard_normal((R, N, 3))
    st, md = rng.integers(0, 3, (R, N, nb)), rng.integers(0, 3, (R, N, nb))
    out = rd.coarse_grain_ensemble(r, t3, st, md, k)
    for i in range(R):
        o = RO.cg_arrays(r[i], t3[i], st[i], md[i], k, k ** (1 / 3))
        for name, want in zip(("r", "t3", "t2", "states", "chemical_mods"), o):
            assert np.array_equal(out[name][i], want), (name, i)


@pytest.mark.parametrize("R,M,n_ref", [(3, 6, 100), (2, 20, 613), (4, 3, 9), (2, 2, 40), (2, 300, 1000), (1, 5, 4000)])
def test_ensemble_refinement_matches_oracle(backend, R, M, n_ref):
    import chromo_b200.util.rediscretize as rd
    rng = np.random.default_rng(R + M + n_ref)
    cg = np.cumsum(rng.standard_normal((R, M, 3)), axis=1) * 30.0
    t = rng.standard_normal((R, M, 3))
    L = RO.refine_layout(M, n_ref)
    np.random.seed(9)
    out = rd.refine_ensemble(cg, t, n_ref, 16.5, confine_length=60.0, seed=None)
    np.random.seed(9)
    xi_r = np.random.standard_normal((R, L["draws"], 3))
    xi_t = np.random.standard_normal((R, L["draws"], 3))
    scaling = (n_ref / M) ** (1 / 3)
    for i in range(R):
        want = RO.confine(RO.refine_path(cg[i], n_ref, 16.5 / scaling, xi_r[i]) * scaling, 60.0)
        assert rel_close(out["r"][i], want)
        t3, t2 = RO.orient(RO.refine_path(t[i], n_ref, np.pi, xi_t[i]))
        assert rel_close(out["t3"][i], t3) and rel_close(out["t2"][i], t2)


def test_device_side_deviates_and_properties(backend):
    """Philox deviates on the device: reproducible per seed, different across seeds
    and replicas; every bridge starts exactly on its coarse bead (B[0] = 0), steps
    average to the requested spacing, orientations are unit and orthogonal."""
    import chromo_b200.util.rediscretize as rd
    rng = np.random.default_rng(5)
    R, M, n_ref, sp = 4, 12, 1003, 40.0
    cg = np.cumsum(rng.standard_normal((R, M, 3)), axis=1) * 60.0
    cg[1] = cg[0]
    a, _, _ = rd._refine(cg, n_ref, sp, orientations=False, out_scale=1.0, seed=11)
    b, _, _ = rd._refine(cg, n_ref, sp, orientations=False, out_scale=1.0, seed=11)
    c, _, _ = rd._refine(cg, n_ref, sp, orientations=False, out_scale=1.0, seed=12)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert not np.array_equal(a[0], a[1])  # same coarse path, different replica stream
    L = RO.refine_layout(M, n_ref)
    starts = L["h1"] + L["seg"] * np.arange(M - 1)
    assert np.array_equal(a[:, starts], cg[:, :M - 1])
    inner = a[:, L["h1"]:L["h1"] + (M - 2) * L["seg"] + 1]
    steps = np.linalg.norm(np.diff(inner, axis=1), axis=2)
    assert abs(steps.mean() / sp - 1) < 0.05
    z = a[:, :L["h1"]]  # free end: unit-direction steps of exactly `sp`
    assert np.allclose(np.linalg.norm(np.diff(z, axis=1), axis=2), sp, rtol=1e-12)
    t3, t2, _ = rd._refine(rng.standard_normal((R, M, 3)), n_ref, np.pi, orientations=True, seed=3)
    assert np.allclose(np.linalg.norm(t3, axis=2), 1, atol=1e-14) and np.allclose(np.linalg.norm(t2, axis=2), 1, atol=1e-14)
    assert np.max(np.abs(np.sum(t3 * t2, axis=2))) < 1e-14
    # the deviates are isotropic: free-end steps are unit vectors with zero mean
    d, _, _ = rd._refine(np.zeros((16, 3, 3)), 8001, 1.0, orientations=False, seed=1)
    w = np.diff(d[:, :2000], axis=1).reshape(-1, 3)
    assert np.allclose(np.linalg.norm(w, axis=1), 1.0, rtol=1e-12) and np.all(np.abs(w.mean(axis=0)) < 0.02)
    assert np.all(np.abs((w * w).mean(axis=0) - 1 / 3) < 0.02)


def test_argument_errors(backend):
    import chromo_b200.util.rediscretize as rd
    from chromo_b200._lib import ChromoError
    with pytest.raises(ZeroDivisionError):
        rd.get_refined_path(np.zeros((5, 3)), 8)       # 2 refined beads per coarse bond
    with pytest.raises(ValueError):
        rd.coarse_grain_ensemble(np.zeros((1, 4, 3)), np.ones((1, 4, 3)), None, None, 5)
    with pytest.raises(ChromoError):
        rd.coarse_grain_ensemble(np.zeros((1, 40, 3)), np.ones((1, 40, 3)), np.full((1, 40, 1), 99), None, 5)
    with pytest.raises(NotImplementedError):
        rd.get_avg_in_intervals(np.zeros((10, 3)), {0: (0, 3), 1: (3, 10)})


# ---- full size on the GPU: size-independent properties --------------------------------------------
@pytest.mark.gpu
def test_c2_ensemble_round_trip_properties(cuda_backend):
    """1,024 replicas x 10,000 beads coarse-grained by 5 and refined back: interval
    means re-aggregate to the global mean, majority states of constant blocks are
    the block value, bridges start on their coarse beads, confinement holds."""
    import chromo_b200.util.rediscretize as rd
    rng = np.random.default_rng(0)
    R, N, k = 1024, 10000, 5
    r = np.cumsum(rng.standard_normal((R, N, 3), dtype=np.float32).astype(np.float64), axis=1) * 16.5
    t3 = rng.standard_normal((R, N, 3), dtype=np.float32).astype(np.float64)
    blocks = rng.integers(0, 3, (R, N // k, 1))
    st = np.repeat(blocks, k, axis=1)
    out = rd.coarse_grain_ensemble(r, t3, st, st, k)
    M = N // k
    assert out["r"].shape == (R, M, 3) and np.array_equal(out["states"], blocks)
    f = k ** (1 / 3)
    assert np.allclose(out["r"].mean(axis=1) * f, r.mean(axis=1), rtol=1e-9, atol=1e-9)
    pick = rng.integers(0, R, 4)
    for i in pick:
        o = RO.cg_arrays(r[i], t3[i], st[i], st[i], k, f)
        assert np.array_equal(out["r"][i], o[0]) and np.array_equal(out["t3"][i], o[1])
        assert np.array_equal(out["t2"][i], o[2])
    assert out["kernel_ms"] > 0
    Rr = 64
    ref = rd.refine_ensemble(out["r"][:Rr], out["t3"][:Rr], N + 1, 16.5, confine_length=0.0, seed=7)
    L = RO.refine_layout(M, N + 1)
    assert ref["r"].shape == (Rr, N + 1, 3)
    starts = L["h1"] + L["seg"] * np.arange(M - 1)
    scaling = ((N + 1) / M) ** (1 / 3)
    assert np.allclose(ref["r"][:, starts], out["r"][:Rr, :M - 1] * scaling, rtol=1e-15, atol=0)
    assert np.allclose(np.linalg.norm(ref["t3"], axis=2), 1, atol=1e-14)


def test_ensemble_pipeline_coarse_grain_mc_refine(backend):
    """The chromosome-scale workflow on a small ensemble: coarse-grain every replica, run MC on the
    coarse chains, refine back with a binding equilibration; mass is conserved at each stage."""
    import math
    import oracle as O
    from chromo_b200.ensemble import ReplicaEnsemble, default_moves
    from chromo_b200.util import poly_paths as paths
    rng = np.random.default_rng(2)
    R, N, k = 2, 240, 4
    Rc = 120.0
    r = paths.confined_gaussian_walk(N, np.full(N - 1, 16.5), "Spherical", Rc, rng, replicas=R)
    t3, t2 = paths.estimate_tangents_from_coordinates(r)
    mods = paths.synthetic_marks(N, 1, rng, replicas=R)
    hp1 = dict(O.HP1)
    grid = dict(x_width=2.4 * Rc, nx=10, y_width=2.4 * Rc, ny=10, z_width=2.4 * Rc, nz=10, confine_type="Spherical",
                confine_length=Rc, vf_limit=0.5)
    ens = ReplicaEnsemble(r, t3, t2, np.zeros((R, N, 1), dtype=np.int64), mods, binders=[hp1],
                          bond_params=O.bond_params(np.full(N - 1, 16.5), 53.0), grid=grid,
                          bead_vol=(4 / 3) * math.pi * 125.0, moves=default_moves(R, N, 16.5), min_spacing=16.5)
    cg = ens.coarse_grained(k)
    assert cg.N == N // k and cg.grid["nx"] == int(round(10 / k ** (1 / 3))) + 1
    want = RO.cg_arrays(r[0], t3[0], ens.states[0], mods[0], k, k ** (1 / 3))
    assert np.array_equal(cg.r[0], want[0]) and np.array_equal(cg.chemical_mods[0], want[4])
    vol = lambda e: e.grid["x_width"] * e.grid["y_width"] * e.grid["z_width"] / (e.grid["nx"] * e.grid["ny"] * e.grid["nz"])
    assert np.allclose(cg.density()[..., 0].sum(axis=1) * vol(cg), cg.N, rtol=1e-9)
    cg.mc_sim(1, 1.0, 5)
    fine = cg.refined(N + 1, 16.5, np.concatenate([mods, mods[:, :1]], axis=1), seed=9, binding_equilibration=100)
    assert fine.N == N + 1 and fine.states.min() >= 0 and fine.states.max() <= 2 and fine.states.sum() > 0
    assert np.all(np.linalg.norm(fine.r, axis=2) <= fine.grid["confine_length"] * 1.02)
    d = fine.density()
    assert np.allclose(d[..., 0].sum(axis=1) * vol(fine), fine.N, rtol=1e-9)
    assert np.allclose(d[..., 1].sum(axis=1) * vol(fine), fine.states[..., 0].sum(axis=1), rtol=1e-9)
    for e in (fine, cg, ens):
        e.close()
