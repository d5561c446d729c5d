"""Snapshot I/O and resume: the polymer CSV schema against files written by
the reference's own writer (tests/golden/snapshot_*.csv, made by make_golden.py from
oracle/_ref), the run-directory layout, and `continue_polymer_in_field_simulation`."""
import os

import numpy as np
import pytest

import oracle as O
from common import GOLDEN, load_golden


def _polymer(name):
    import chromo_b200.polymers as ply
    spec, _ = load_golden(f"{name}_spec")
    N, nb = spec["N"], spec["nb"]
    kw = dict(bead_length=spec["bead_length"], bead_rad=float(spec["bead_rad"]), t3=spec["t3"].copy(),
              t2=spec["t2"].copy(), states=spec["states"].reshape(N, nb).copy(),
              binder_names=np.array([b["name"] for b in spec["binders"]]),
              chemical_mods=spec["mods"].reshape(N, nb).copy(),
              chemical_mod_names=np.array([f"mod{j}" for j in range(nb)]), max_binders=spec["max_binders"])
    if spec["lp"] == 53.0:
        return ply.Chromatin("Chr-1", spec["r"].copy(), **kw), ply.Chromatin
    return ply.SSWLC("homopolymer", spec["r"].copy(), lp=float(spec["lp"]), **kw), ply.SSWLC


@pytest.mark.parametrize("name", ["snapshot_chromatin", "snapshot_sswlc"])
def test_csv_is_byte_identical_to_the_reference_writer(name, tmp_path):
    p, _ = _polymer(name)
    out = tmp_path / "p.csv"
    p.to_csv(str(out))
    assert out.read_bytes() == (GOLDEN / f"{name}.csv").read_bytes()
    p.to_file(str(tmp_path / "q.csv"))                       # synonym
    assert (tmp_path / "q.csv").read_bytes() == out.read_bytes()


@pytest.mark.parametrize("name", ["snapshot_chromatin", "snapshot_sswlc"])
def test_reads_the_reference_csv(name):
    p, cls = _polymer(name)
    q = cls.from_file(str(GOLDEN / f"{name}.csv"))
    assert q.name == name                                     # named after the file (polymers.pyx:788-789)
    for a in ("r", "t3", "t2", "states", "chemical_mods", "bead_length"):
        assert np.array_equal(getattr(p, a), getattr(q, a)), a
    assert q.states.dtype == np.int64 and q.r.flags["C_CONTIGUOUS"]
    assert list(q.binder_names) == list(p.binder_names)
    assert list(q.chemical_mod_names) == list(p.chemical_mod_names)
    assert q.max_binders == p.max_binders and float(q.lp) == float(p.lp)
    assert cls.from_file(str(GOLDEN / f"{name}.csv"), name="other").name == "other"
    # elastic parameters follow from bead_length / lp, so they survive the round trip
    assert np.array_equal(p.eps_bend, q.eps_bend) and np.array_equal(p.gamma, q.gamma)


def test_run_folder_helpers(tmp_path):
    from chromo_b200.util import poly_stat, reproducibility as rp
    a = rp.get_unique_subfolder(tmp_path / "sim_")
    b = rp.get_unique_subfolder(tmp_path / "sim_")
    assert (a.name, b.name) == ("sim_1", "sim_2") and (b / "acceptance_trackers").is_dir()
    assert rp.get_unique_subfolder_name(tmp_path / "sim_").name == "sim_3"
    for k in range(9, 12):
        (tmp_path / f"sim_{k}").mkdir()
    assert poly_stat.get_latest_simulation(str(tmp_path)) == "sim_11"     # numeric, not lexicographic
    for f in ("Chr-1", "Chr-2", "Chr-1-0.csv", "Chr-1-2.csv", "Chr-1-10.csv", "Chr-2-0.csv", "notes.txt"):
        (a / f).write_text("")
    assert poly_stat.find_polymers_in_output_dir(str(a)) == ["Chr-1", "Chr-2"]
    assert poly_stat.get_latest_configuration("Chr-1", str(a)).endswith("/Chr-1-10.csv")
    with pytest.raises(FileNotFoundError):
        poly_stat.get_latest_configuration("Chr-3", str(a))


def test_snapshots_and_resume(backend, tmp_path):
    """polymer_in_field writes the reference's run layout; continuing loads the latest
    snapshot and carries on in a new sim_<k> folder with consecutive snapshot numbers."""
    import chromo_b200.binders as bnd
    import chromo_b200.fields as fld
    import chromo_b200.polymers as ply
    from chromo_b200.mc import continue_polymer_in_field_simulation, get_amplitude_bounds, polymer_in_field
    spec = O.make_spec(N=60, nb=1, seed=21, random_states=False)
    f = spec["field"]

    def make(name="Chr-1"):
        hp1 = bnd.get_by_name("HP1")
        binders = bnd.make_binder_collection([hp1])
        p = ply.Chromatin(name, spec["r"].copy(), bead_length=spec["bead_length"], t3=spec["t3"].copy(),
                          t2=spec["t2"].copy(), states=spec["states"].copy(), binder_names=np.array(["HP1"]),
                          chemical_mods=spec["mods"].copy(), chemical_mod_names=np.array(["H3K9me3"]))
        field = fld.UniformDensityField([p], binders, f["x_width"], f["nx"], f["y_width"], f["ny"], f["z_width"],
                                        f["nz"], confine_type=f["confine_type"],
                                        confine_length=f["confine_length"], chi=f["chi"])
        return p, binders, field

    p, binders, field = make()
    bb, mb = get_amplitude_bounds([p])
    out = str(tmp_path / "output")
    polys = polymer_in_field([p], binders, field, 2, 3, bb, mb, random_seed=5, output_dir=out)
    assert polys[0] is p
    run = tmp_path / "output" / "sim_1"
    assert sorted(os.listdir(run)) == ["Chr-1", "Chr-1-0.csv", "Chr-1-1.csv", "Chr-1-2.csv", "acceptance_trackers"]
    assert len(os.listdir(run / "acceptance_trackers")) == 5 * 3        # one log per move type and snapshot
    assert p.log_path == f"{run}/Chr-1_config_log.csv"
    start = ply.Chromatin.from_file(str(run / "Chr-1"))
    assert np.array_equal(start.r, spec["r"])                          # the initial configuration
    last = ply.Chromatin.from_file(str(run / "Chr-1-2.csv"))
    assert np.array_equal(last.r, p.r) and np.array_equal(last.states, p.states)
    assert np.array_equal(last.t3, p.t3) and np.array_equal(last.t2, p.t2)  # shortest round-trip floats

    # resume: a fresh field object, the polymer comes from the latest snapshot
    _, binders2, field2 = make()
    polys2 = continue_polymer_in_field_simulation(ply.Chromatin, binders2, field2, out, 2, 2, random_seed=6)
    run2 = tmp_path / "output" / "sim_2"
    assert sorted(os.listdir(run2)) == ["Chr-1", "Chr-1-3.csv", "Chr-1-4.csv", "acceptance_trackers"]
    assert polys2[0].name == "Chr-1"
    resumed_from = ply.Chromatin.from_file(str(run2 / "Chr-1"))
    assert np.array_equal(resumed_from.r, last.r) and np.array_equal(resumed_from.states, last.states)
    assert not np.array_equal(polys2[0].r, last.r)                       # it moved on
    # the field the continuation ran in was rebuilt from the loaded configuration: its density
    # equals a recompute from the final configuration
    d = field2.density.copy()
    field2.update_all_densities_for_all_polymers()
    assert np.allclose(d, field2.density, rtol=1e-9, atol=1e-15)


def test_ensemble_snapshot_round_trip(backend, tmp_path):
    """One .npz per replica batch: reloading restores positions, states, controller state and
    the derived densities exactly; any replica exports to the reference's CSV schema."""
    import chromo_b200.polymers as ply
    from chromo_b200.ensemble import ReplicaEnsemble, default_moves
    R, N = 3, 80
    specs = [O.make_spec(N=N, nb=2, seed=30 + i, cross_talk=-1.0) for i in range(R)]
    st = lambda k: np.stack([s[k] for s in specs])
    bond = O.bond_params(specs[0]["bead_length"], 53.0)
    kw = dict(binders=specs[0]["binders"], bond_params=bond, grid=specs[0]["field"],
              bead_vol=(4 / 3) * np.pi * 5.0 ** 3, chi=[0.5, 1.0, 1.5], moves=default_moves(R, N, 16.5))
    ens = ReplicaEnsemble(st("r"), st("t3"), st("t2"), st("states"), st("mods"), **kw)
    ens.mc_sim(3, 1.0, 17)
    ens.sync()
    path = tmp_path / "batch_0.npz"
    ens.save_snapshot(path)
    dens = ens.density().copy()
    moves = ens.moves.copy()
    ens2 = ReplicaEnsemble(st("r"), st("t3"), st("t2"), st("states"), st("mods"),
                           **dict(kw, chi=1.0, moves=default_moves(R, N, 16.5)))
    ens2.load_snapshot(path)
    for a in ("r", "t3", "t2", "states", "chemical_mods", "chi", "mu"):
        assert np.array_equal(getattr(ens, a), getattr(ens2, a)), a
    assert ens2.moves.tobytes() == moves.tobytes()
    assert np.allclose(ens2.density(), dens, rtol=1e-9, atol=1e-15)  # incremental updates are fixed point (quantum ~1e-17, mc_kernel.cuh fx_format)
    # any replica of the batch exports to the reference's CSV schema
    csv = tmp_path / "Chr-2-0.csv"
    ens2.replica_to_csv(1, csv, bead_length=specs[1]["bead_length"])
    q = ply.Chromatin.from_file(str(csv))
    assert np.array_equal(q.r, ens.r[1]) and np.array_equal(q.states, ens.states[1])
    assert list(q.binder_names) == ["HP1", "PRC1"]
    with pytest.raises(ValueError):
        small = ReplicaEnsemble(st("r")[:2], st("t3")[:2], st("t2")[:2], st("states")[:2], st("mods")[:2],
                                **dict(kw, chi=1.0, moves=default_moves(2, N, 16.5)))
        small.load_snapshot(path)
    ens.close()
    ens2.close()
