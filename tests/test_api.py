"""The reference-facing Python API (chromo_b200.polymers / binders / fields /
mc) driven exactly like a chromo script, on both backends, against golden
vectors from the reference and -- where oracle/_ref is built -- against the
live reference."""
import numpy as np
import pytest

import oracle as O
from common import close, load_golden


def build(spec, lazy=False):
    import chromo_b200.binders as bnd
    import chromo_b200.fields as fld
    import chromo_b200.polymers as ply
    objs = []
    for b in spec["binders"]:
        o = bnd.get_by_name(b["name"])
        for k in ("sites_per_bead", "bind_energy_mod", "bind_energy_no_mod", "interaction_energy",
                  "chemical_potential", "interaction_radius"):
            setattr(o, k, b[k])
        o.interaction_volume = (4.0 / 3.0) * np.pi * b["interaction_radius"] ** 3
        o.cross_talk_interaction_energy = dict(b["cross_talk"])
        o.cross_talk_field_energy_prefactor = {}
        objs.append(o)
    binders = bnd.make_binder_collection(objs)
    N, nb = spec["N"], spec["nb"]
    kw = dict(bead_length=spec["bead_length"], t3=spec["t3"].copy(),
              t2=spec["t2"].copy(), states=spec["states"].reshape(N, nb).copy(),
              binder_names=np.array([b["name"] for b in spec["binders"]]),
              chemical_mods=spec["mods"].reshape(N, nb).copy(),
              chemical_mod_names=np.array([f"m{j}" for j in range(nb)]))
    if spec.get("bp_wrap") is not None:  # detailed nucleosomes: polymers.pyx:2455
        cls = ply.DetailedChromatin2 if spec.get("no_diameter") else ply.DetailedChromatin
        p = cls("c", spec["r"].copy(), bp_wrap=spec["bp_wrap"], lp=spec["lp"], lt=spec["lt"], **kw)
    elif spec.get("lt") is not None:  # twist: polymers.pyx:1889
        p = ply.SSTWLC("c", spec["r"].copy(), lp=spec["lp"], lt=spec["lt"], **kw)
    else:
        p = ply.Chromatin("c", spec["r"].copy(), **kw)
    f = spec["field"]
    field = fld.UniformDensityField([p], binders, f["x_width"], f["nx"], f["y_width"], f["ny"], f["z_width"],
                                    f["nz"], confine_type=f["confine_type"], confine_length=f["confine_length"],
                                    chi=f["chi"], vf_limit=f["vf_limit"],
                                    assume_fully_accessible=f.get("assume_fully_accessible", 1),
                                    fast_field=f.get("fast_field", 0), n_points=f.get("n_points", 1000))
    return p, binders, field


@pytest.mark.parametrize("name", ["static_c2", "static_c3", "static_tw2", "static_av", "static_dc"])
def test_construction_and_energies(backend, name):
    spec, g = load_golden(name)
    p, binders, field = build(spec)
    for k in ("eps_bend", "eps_par", "eps_perp", "gamma", "eta"):
        assert np.array_equal(getattr(p, k), g[k]), k
    assert np.allclose(field.density, g["density"], rtol=1e-12, atol=0)
    assert np.array_equal(field.density != 0, g["density"] != 0)
    assert field.vol_bin == float(g["vol_bin"]) and p.beads[0].vol == float(g["bead_vol"])
    bd = field.binder_dict
    assert np.array_equal([b["field_energy_prefactor"] for b in bd], g["field_pref"])
    assert np.array_equal([b["interaction_energy_intranucleosome"] for b in bd], g["e_intra"])
    assert close(field.compute_E(p), float(g["E_field"]))
    assert close(p.compute_E(), float(g["E_poly"]))
    if "lt" in spec:  # SSTWLC: the twist term is separable (compute_E_no_twist polymers.pyx:2287-2319)
        twist = p.compute_E() - p.compute_E_no_twist()
        assert twist > 0 and close(p.compute_E(), float(g["E_poly"]))  # the term is restored afterwards
        with pytest.raises(Exception, match="Twist modulus must be positive"):
            p._polymer_engine().set_twist_params(np.zeros(spec["N"] - 1), p.natural_twist)


@pytest.mark.parametrize("name", ["mcsim_c2", "mcsim_c3", "mcsim_tw", "mcsim_ff", "mcsim_dc", "mcsim_dc3", "mcsim_order"])
def test_mc_sim_drop_in(backend, name):
    """all_moves + SimpleControl + mc_sim, replaying the reference's RNG streams."""
    from chromo_b200.mc import get_amplitude_bounds, mc_controller as ctrl, set_rng_mode
    from chromo_b200.mc.mc_sim import mc_sim
    spec, g = load_golden(name)
    p, binders, field = build(spec)
    bb, mb = get_amplitude_bounds([p])
    cs = ctrl.all_moves("/tmp/chromo_b200_test", bb.bounds, mb.bounds, ctrl.SimpleControl)
    for c, k in zip(cs, g["per_cycle"]):
        c.move.num_per_cycle = int(k)
    set_rng_mode("replay")
    try:
        field._engine_for(p).srand(int(g["srand_seed"]))
        run = [cs[i] for i in g["order"]] if "order" in g else cs  # a controller list in the caller's own order
        mc_sim([p], binders, int(g["steps"]), run, field, float(g["mu_adjust"]), int(g["np_seed"]))
    finally:
        set_rng_mode("philox")
    tol = 0 if backend == "emu" else 1e-7
    assert np.allclose(p.r, g["final_r"], rtol=0, atol=tol)
    assert np.allclose(p.t3, g["final_t3"], rtol=0, atol=tol)
    assert np.array_equal(p.states, g["final_states"])
    assert np.array_equal(p.r, p.r_trial) and np.array_equal(p.states, p.states_trial)
    assert [c.move.num_attempt for c in cs] == list(g["num_attempt"])
    assert [c.move.num_success for c in cs] == list(g["num_success"])
    assert [c.move.amp_bead for c in cs] == list(g["amp_bead"])
    assert np.array_equal([c.move.amp_move for c in cs], g["amp_move"])
    assert np.array_equal([c.move.acceptance_tracker.acceptance_rate for c in cs], g["acceptance_rate"])
    vol_bin = field.vol_bin
    assert np.allclose(field.density, g["final_density"], rtol=1e-9, atol=1e-9 / vol_bin)


def test_null_field_confinement_and_polymer_in_field(backend):
    """Tutorial-2 style: SSWLC homopolymer, NullField with a spherical confinement,
    physical moves only, production RNG, snapshot driver."""
    import chromo_b200.binders as bnd
    import chromo_b200.fields as fld
    import chromo_b200.polymers as ply
    from chromo_b200.mc import get_amplitude_bounds, mc_controller as ctrl, polymer_in_field
    from chromo_b200.util import mu_schedules
    rng = np.random.default_rng(5)
    N, Rc = 120, 60.0
    r = O.confined_walk(N, 5.0, Rc, rng)
    t3, t2 = O.tangents_from_coords(r, rng)
    p = ply.SSWLC("homopolymer", r, bead_length=np.full(N - 1, 5.0), lp=10.0, t3=t3, t2=t2)
    assert p.num_binders == 1 and p.binder_names[0] == "null_reader"
    binders = bnd.make_binder_collection([bnd.get_by_name("null_reader")])
    field = fld.NullField([p], confine_type="Spherical", confine_length=Rc)
    bb, mb = get_amplitude_bounds([p])
    cs = ctrl.all_moves_except_binding_state("/tmp/chromo_b200_test", bb.bounds, mb.bounds, ctrl.SimpleControl)
    E0 = p.compute_E()
    snaps = []
    polymer_in_field([p], binders, field, 5, 3, bb, mb, mc_move_controllers=cs, random_seed=3,
                     mu_schedule=mu_schedules.Schedule(mu_schedules.linear_step_for_negative_cp),
                     save_snapshot=lambda i, polys, f, c: snaps.append(polys[0].r.copy()))
    assert len(snaps) == 3 and not np.array_equal(snaps[0], snaps[2])
    assert [c.move.num_attempt for c in cs] == [75, 75, 75, 75]
    assert np.all(np.linalg.norm(p.r, axis=1) <= Rc + 1e-9)          # the confinement held
    assert np.allclose(np.linalg.norm(p.t3, axis=1), 1.0, atol=1e-9)  # rotations are orthogonal
    assert p.compute_E() != E0
    # the same run again is identical (Philox keyed by the seeds)
    p2 = ply.SSWLC("homopolymer", r, bead_length=np.full(N - 1, 5.0), lp=10.0, t3=t3, t2=t2)
    f2 = fld.NullField([p2], confine_type="Spherical", confine_length=Rc)
    cs2 = ctrl.all_moves_except_binding_state("/tmp/chromo_b200_test", bb.bounds, mb.bounds, ctrl.SimpleControl)
    polymer_in_field([p2], binders, f2, 5, 3, bb, mb, mc_move_controllers=cs2, random_seed=3,
                     mu_schedule=mu_schedules.Schedule(mu_schedules.linear_step_for_negative_cp))
    assert np.array_equal(p.r, p2.r)


def test_move_functions_and_mc_step(backend):
    """Proposals through the move functions leave the trial state in the polymer,
    like the reference's cpdef move functions; mc_step runs one attempt."""
    from chromo_b200.mc import mc_controller as ctrl, get_amplitude_bounds, move_funcs as mf
    from chromo_b200.mc.mc_sim import mc_step
    spec, g = load_golden("static_c2")
    p, binders, field = build(spec)
    inds = mf.slide(p, 2.0, 10)
    assert len(inds) >= 1 and np.array_equal(np.diff(inds), np.ones(len(inds) - 1))
    d = p.r_trial[inds] - p.r[inds]
    assert np.allclose(d, d[0]) and 0 < np.linalg.norm(d[0]) <= 2.0 + 1e-12   # one rigid translation
    assert np.array_equal(p.t3_trial[inds], p.t3[inds])
    inds = mf.crank_shaft(p, 0.3, 30)
    a = np.linalg.norm(p.r_trial[inds][1:] - p.r_trial[inds][:-1], axis=1) if len(inds) > 1 else np.zeros(0)
    b = np.linalg.norm(p.r[inds][1:] - p.r[inds][:-1], axis=1) if len(inds) > 1 else np.zeros(0)
    assert np.allclose(a, b)                                                   # a rigid rotation
    bb, mb = get_amplitude_bounds([p])
    cs = ctrl.all_moves("/tmp/chromo_b200_test", bb.bounds, mb.bounds, ctrl.SimpleControl)
    r0 = p.r.copy()
    for _ in range(6):
        mc_step(cs[2].move, p, binders, field, True, False)
    assert cs[2].move.num_attempt == 6 and 0 <= cs[2].move.num_success <= 6
    assert (cs[2].move.num_success > 0) == (not np.array_equal(p.r, r0))


def test_argument_validation(backend):
    import chromo_b200.binders as bnd
    import chromo_b200.fields as fld
    import chromo_b200.polymers as ply
    spec, g = load_golden("static_c2")
    N = spec["N"]
    with pytest.raises(ValueError, match="chemical state must be given a name"):
        ply.Chromatin("c", spec["r"], bead_length=spec["bead_length"], t3=spec["t3"], t2=spec["t2"],
                      states=np.zeros((N, 2), dtype=np.int64), binder_names=np.array(["HP1"]),
                      chemical_mods=np.zeros((N, 2), dtype=np.int64), chemical_mod_names=np.array(["a", "b"]))
    with pytest.raises(TypeError):
        ply.Chromatin("c", spec["r"].tolist(), bead_length=spec["bead_length"])
    with pytest.raises(ValueError, match="No binders found"):
        bnd.get_by_name("HP2")
    p, binders, field = build(spec)
    with pytest.raises(NotImplementedError, match="same binders"):
        fld.UniformDensityField([p], bnd.make_binder_collection([bnd.hp1, bnd.prc1]), 10, 2, 10, 2, 10, 2)
    with pytest.raises(ValueError, match="Confinement type"):
        f = spec["field"]
        fld.UniformDensityField([p], binders, f["x_width"], f["nx"], f["y_width"], f["ny"], f["z_width"],
                                f["nz"], confine_type="Ellipsoid", confine_length=1.0)


def test_field_csv_round_trip(backend, tmp_path):
    """tests/test_fields.py:12-59 of the reference: to_file / from_file."""
    import chromo_b200.fields as fld
    spec, g = load_golden("static_c2")
    p, binders, field = build(spec)
    path = tmp_path / "UniformDensityField"
    field.to_file(path)
    f2 = fld.UniformDensityField.from_file(path, [p], binders)
    assert f2 == field and np.allclose(f2.density, field.density, rtol=1e-12)
    with pytest.raises(ValueError, match="polymers, but"):
        fld.UniformDensityField.from_file(path, [p, p], binders)


def test_accessible_volumes(backend):
    """assume_fully_accessible=0: voxels cut by the sphere get a reduced volume
    (fields.pyx:714-951) and the kernels divide by it."""
    import chromo_b200.fields as fld
    spec, g = load_golden("static_c2")
    p, binders, _ = build(spec)
    f = spec["field"]
    field = fld.UniformDensityField([p], binders, f["x_width"], f["nx"], f["y_width"], f["ny"], f["z_width"],
                                    f["nz"], confine_type="Spherical", confine_length=f["confine_length"],
                                    chi=f["chi"], assume_fully_accessible=0)
    av = field.access_vols
    assert av.min() < field.vol_bin and av.max() == field.vol_bin
    # mass conservation with per-voxel volumes: sum rho * V_access = N
    assert abs((field.density[:, 0] * av).sum() - spec["N"]) < 1e-9 * spec["N"]


@pytest.mark.ref
def test_against_live_reference(backend):
    """Random problems beyond the goldens: the same script on the reference's
    own Cython build (oracle/_ref) and on chromo_b200, replayed RNG."""
    if not O.ref_available():
        pytest.skip("oracle/_ref not built on this box")
    from chromo_b200.mc import get_amplitude_bounds, mc_controller as ctrl, set_rng_mode
    from chromo_b200.mc.mc_sim import mc_sim
    for seed, nb in ((21, 1), (22, 2)):
        spec = O.make_spec(N=180, nb=nb, seed=seed, cross_talk=-0.8 if nb == 2 else 0.0)
        rp, rdf, rfield, M = O.ref_objects(spec)
        rbb, rmb = M["mc"].get_amplitude_bounds([rp])
        rcs = M["mc_controller"].all_moves("/tmp/x", rbb.bounds, rmb.bounds, M["mc_controller"].SimpleControl)
        M["shim"].c_srand(seed)
        with np.errstate(over="ignore"):
            M["mc_sim"].mc_sim([rp], rdf, 6, rcs, rfield, 0.9, seed + 1)
        p, binders, field = build(spec)
        bb, mb = get_amplitude_bounds([p])
        cs = ctrl.all_moves("/tmp/x", bb.bounds, mb.bounds, ctrl.SimpleControl)
        set_rng_mode("replay")
        try:
            field._engine_for(p).srand(seed)
            mc_sim([p], binders, 6, cs, field, 0.9, seed + 1)
        finally:
            set_rng_mode("philox")
        assert [c.move.num_success for c in cs] == [c.move.num_success for c in rcs]
        tol = 0 if backend == "emu" else 1e-7
        assert np.allclose(p.r, np.asarray(rp.r), rtol=0, atol=tol)
        assert np.array_equal(p.states, np.asarray(rp.states))
        assert close(field.compute_E(p), rfield.compute_E(rp), 1e-9 if backend == "emu" else 1e-7)


def test_nucleosome_constants_match_the_pinned_oracle():
    """util.nucleo_geom (product) against the oracle's restatement, which the *_dc goldens pin to the reference."""
    from chromo_b200.util.nucleo_geom import nucleosome_constants
    for bp in (147, 146.5, 127, 1):
        c, o = nucleosome_constants(bp), O.nucleosome_constants(bp)
        want = np.concatenate([o["t3_local"], o["t2_local"], o["r_enter_unit"], [o["r_enter_norm"]], o["r_exit_unit"],
                               [o["r_exit_norm"]], o["a3"], o["a1"]])
        assert np.array_equal(c, want), bp


def test_ensemble_from_detailed_chromatin(backend):
    """ReplicaEnsemble.from_polymers on DetailedChromatin replicas keeps the entry / exit geometry: a crank-shaft
    production mc_sim of the batched engine is the oracle's DetailedChromatin walk on the same Philox streams
    (a plain SSTWLC would accept different moves)."""
    from chromo_b200.ensemble import ReplicaEnsemble
    spec, g = load_golden("static_dc")
    built = [build(spec) for _ in range(2)]
    ens = ReplicaEnsemble.from_polymers([b[0] for b in built], [b[2] for b in built])
    assert np.array_equal(ens.bond_params["nucleosome_constants"], built[0][0].nucleosome_constants)
    E = ens.elastic_energy()
    assert close(E[0], float(g["E_poly"])) and E[0] == E[1]
    mv0 = ens.moves.copy()
    ens.mc_sim(2, 1.0, 77)
    finals = {}
    for detailed in (True, False):
        o = O.OracleSim(spec if detailed else {k: v for k, v in spec.items() if k != "bp_wrap"})
        o.use_production_streams(77, 1, 0)
        omv = O.make_moves(spec["N"], float(np.min(spec["bead_length"])))
        for i in range(5):
            for f in ("amp_move", "amp_bead", "num_attempt", "num_success", "acceptance_rate"):
                setattr(omv[i], f, type(getattr(omv[i], f))(mv0[f][1, i]))
        o.mc_sim(omv, 2, 0)
        finals[detailed] = (o.r.copy(), [m.num_success for m in omv])
    assert [int(x) for x in ens.moves["num_success"][1]] == finals[True][1]
    assert np.allclose(ens.r[1], finals[True][0], rtol=0, atol=1e-7)
    assert not np.allclose(ens.r[1], finals[False][0], rtol=0, atol=1e-3)  # the geometry mattered
    ens.close()


def test_ensemble_from_twisted_polymers_keeps_the_twist_term(backend):
    """ReplicaEnsemble.from_polymers on SSTWLC replicas: the batched elastic energy is each polymer's
    own compute_E (twist included), and a short mc_sim runs the twist kernels."""
    from chromo_b200.ensemble import ReplicaEnsemble
    spec, g = load_golden("static_tw")
    built = [build(spec) for _ in range(2)]
    polys, fields = [b[0] for b in built], [b[2] for b in built]
    polys[1].t2[...] = np.cross(polys[1].t3, [0.0, 0.0, 1.0])
    polys[1].t2[...] /= np.linalg.norm(polys[1].t2, axis=1)[:, None]
    ens = ReplicaEnsemble.from_polymers(polys, fields)
    E = ens.elastic_energy()
    assert close(E[0], float(g["E_poly"])) and close(E[0], polys[0].compute_E())
    assert close(E[1], polys[1].compute_E()) and not close(E[0], E[1])
    ens.mc_sim(1, 1.0, 3)
    assert np.isfinite(ens.elastic_energy()).all() and ens.acceptance()["crank_shaft"] > 0
    ens.close()


def test_ensemble_from_polymers_coarse_grains_and_refines(backend):
    """from_polymers -> coarse_grained -> refined (the workflow INTEGRATION.md advertises): the ensemble carries
    the binders' interaction parameters, so the prefactors are rebuilt on the coarser / finer grid, and the
    field's accessible-volume setting with them."""
    from chromo_b200.ensemble import ReplicaEnsemble
    spec, g = load_golden("static_av2")  # two binders with cross-talk, assume_fully_accessible = 0
    built = [build(spec) for _ in range(2)]
    polys, fields = [b[0] for b in built], [b[2] for b in built]
    ens = ReplicaEnsemble.from_polymers(polys, fields)
    assert ens.assume_fully_accessible == 0
    pre = ens.prefactors_from_binders(ens.binders, ens.grid)
    assert np.allclose(pre[0], g["field_pref"]) and np.allclose(pre[1], g["e_intra"]) and np.allclose(pre[2], g["xpref"])
    cg = ens.coarse_grained(4)
    assert cg.N == spec["N"] // 4 and cg.assume_fully_accessible == 0
    E = cg.field_energy()
    assert np.isfinite(E).all() and E[0] == E[1]
    cg.mc_sim(1, 1.0, 3)
    fine = cg.refined(spec["N"] + 1, 16.5, np.concatenate([spec["mods"], spec["mods"][:1]])[None], seed=4)
    assert fine.N == spec["N"] + 1 and fine.assume_fully_accessible == 0 and np.isfinite(fine.field_energy()).all()
    for e in (fine, cg, ens):
        e.close()



def test_ensemble_page_locks_its_host_arrays(backend):
    """pin_host: the ensemble registers its r / t3 / t2 / states / chemical_mods buffers (chromo_host_register)
    for the host-in / host-out path and releases them on close; results do not depend on it."""
    from chromo_b200.ensemble import ReplicaEnsemble, default_moves
    spec = O.make_spec(N=80, nb=1, seed=5)
    st = lambda k: np.stack([spec[k]] * 3)
    out = []
    for pin in (True, False):
        ens = ReplicaEnsemble(st("r"), st("t3"), st("t2"), st("states"), st("mods"), binders=spec["binders"],
                              bond_params=O.bond_params(spec["bead_length"], 53.0), grid=spec["field"],
                              bead_vol=(4 / 3) * np.pi * 125.0, moves=default_moves(3, 80, 16.5), pin_host=pin)
        assert len(ens._pinned) == (5 if pin else 0)
        ens.mc_sim(2, 1.0, 9)
        out.append((ens.r.copy(), ens.states.copy()))
        ens.close()
        assert ens._pinned == []
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.gpu
def test_host_register_on_the_gpu():
    """cudaHostRegister for real: a numpy array is locked once (a second registration is refused), memory torch
    already pinned is reported as such."""
    import torch
    from chromo_b200.engine import host_register, host_unregister
    a = np.zeros(1 << 20)
    assert host_register(a) is True
    assert host_register(a) is False          # already page-locked: nothing to do
    host_unregister(a)
    t = torch.empty(1 << 20, dtype=torch.float64, pin_memory=True)
    assert host_register(t.numpy()) is False
    with pytest.raises(Exception):
        host_unregister(np.zeros(16))         # never registered
