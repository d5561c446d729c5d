#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE ITSELF (oracle/_ref, the
unmodified Cython build of /root/reference/chromo made by oracle/build_ref.py).

Run in the authoring container only (needs /root/reference to have been built
into oracle/_ref):   python tests/golden/make_golden.py
Outputs tests/golden/*.npz -- small, committed, consumed by tests/ on any box.

Each file stores the full input `spec` (arrays + a JSON blob for binder/field
scalars) next to the reference's outputs, so the tests need nothing but numpy.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent


def spec_to_npz(spec):
    meta = dict(N=spec["N"], nb=spec["nb"], lp=spec["lp"], bead_rad=spec["bead_rad"],
                binders=spec["binders"], field=spec["field"], max_binders=spec["max_binders"])
    if spec.get("lt") is not None:
        meta["lt"] = spec["lt"]
    if spec.get("bp_wrap") is not None:
        meta["bp_wrap"] = spec["bp_wrap"]
    if spec.get("no_diameter"):
        meta["no_diameter"] = 1
    arrs = {k: np.asarray(spec[k]) for k in ("r", "t3", "t2", "states", "mods", "bead_length")}
    arrs["meta"] = np.array(json.dumps(meta))
    return arrs


def golden_static(name, spec):
    """Densities, total energies and derived parameters right after construction."""
    poly, df, field, M = O.ref_objects(spec)
    out = spec_to_npz(spec)
    out["density"] = np.asarray(field.density).copy()
    out["E_field"] = field.compute_E(poly)
    out["density_after_E"] = np.asarray(field.density).copy()
    out["E_poly"] = poly.compute_E()
    for k in ("eps_bend", "eps_par", "eps_perp", "gamma", "eta"):
        out[k] = np.asarray(getattr(poly, k)).copy()
    bd = field.binder_dict
    out["field_pref"] = np.array([b["field_energy_prefactor"] for b in bd])
    out["e_intra"] = np.array([b["interaction_energy_intranucleosome"] for b in bd])
    out["xpref"] = np.array([[b["cross_talk_field_energy_prefactor"][c["name"]] for c in bd] for b in bd],
                            dtype=float)
    out["vol_bin"] = field.vol_bin
    out["bead_vol"] = poly.beads[0].vol
    if spec["field"].get("assume_fully_accessible", 1) == 0:
        av = field.access_vols  # {bin: volume} (fields.pyx:714-770)
        out["access_vols"] = np.array([av[i] for i in range(field.n_bins)], dtype=float)
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, "E_field", out["E_field"], "E_poly", out["E_poly"])


def golden_moves(name, spec, nmoves, seed):
    """A chain of single moves: proposal, both dE terms, touched bins,
    density_trial rows, then accept/reject decided by a recorded coin."""
    poly, df, field, M = O.ref_objects(spec)
    sh, ctrl, mc = M["shim"], M["mc_controller"], M["mc"]
    bb, mb = mc.get_amplitude_bounds([poly])
    cs = ctrl.all_moves("/tmp/golden", bb.bounds, mb.bounds, ctrl.SimpleControl)
    rng = np.random.default_rng(seed)
    sh.c_srand(seed)
    np.random.seed(seed)
    rec = dict(move=[], amp_move=[], amp_bead=[], n=[], ind0=[], dE_poly=[], dE_field=[], accept=[],
               n_touched=[])
    inds_all, touched_all, dtrial_all, trial_rows = [], [], [], []
    for it in range(nmoves):
        m = int(rng.integers(0, 5))
        # bounded amplitudes: 0.5x .. 3x the controller's lower bound
        amp_move = float(mb.bounds[O.MOVE_NAMES[m]][0] * rng.uniform(0.5, 3))
        amp_bead = int(rng.integers(1, 40))
        cs[m].move.amp_move, cs[m].move.amp_bead = amp_move, amp_bead
        inds = np.asarray(sh.propose(cs[m].move, poly)).copy()
        n = len(inds)
        nm = O.MOVE_NAMES[m]
        dEp = sh.poly_dE(poly, nm, inds, n)
        dEf = 0.0
        touched = np.zeros(0, dtype=np.int64)
        dtr = np.zeros((0, spec["nb"] + 1))
        if m != 3:
            dEf = sh.field_dE(field, poly, inds, n, m == 4)
            touched = np.nonzero(np.asarray(field.affected_bins_last_move))[0]
            dtr = np.asarray(field.density_trial)[touched].copy()
        rows = np.concatenate([np.asarray(poly.r_trial)[inds], np.asarray(poly.t3_trial)[inds],
                               np.asarray(poly.t2_trial)[inds],
                               np.asarray(poly.states_trial)[inds].astype(float)], axis=1)
        acc = bool(rng.uniform() < np.exp(-(dEp + dEf)))
        if acc:
            cs[m].move.accept(poly, dEp + dEf, inds, n, False, False, False)
            if m != 3:
                sh.commit(field)
        else:
            cs[m].move.reject(poly, dEp + dEf, inds, n, False, False, False)
        for k, v in zip(("move", "amp_move", "amp_bead", "n", "ind0", "dE_poly", "dE_field", "accept",
                         "n_touched"),
                        (m, amp_move, amp_bead, n, inds[0], dEp, dEf, acc, len(touched))):
            rec[k].append(v)
        inds_all.append(inds)
        touched_all.append(touched)
        dtrial_all.append(dtr)
        trial_rows.append(rows)
    out = spec_to_npz(spec)
    out.update({k: np.array(v) for k, v in rec.items()})
    out["inds"] = np.concatenate(inds_all)
    out["touched"] = np.concatenate(touched_all)
    out["dtrial"] = np.concatenate(dtrial_all)
    out["trial_rows"] = np.concatenate(trial_rows)
    out["seed"] = seed
    out["final_r"] = np.asarray(poly.r).copy()
    out["final_t3"] = np.asarray(poly.t3).copy()
    out["final_t2"] = np.asarray(poly.t2).copy()
    out["final_states"] = np.asarray(poly.states).copy()
    out["final_density"] = np.asarray(field.density).copy()
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, "moves", nmoves, "accepted", int(np.sum(rec["accept"])))


def golden_mc_sim(name, spec, steps, srand_seed, np_seed, mu_adjust=1.0, per_cycle=None, order=None):
    """A whole mc_sim call under pinned libc/numpy seeds."""
    poly, df, field, M = O.ref_objects(spec)
    sh, ctrl, mc, mcs = M["shim"], M["mc_controller"], M["mc"], M["mc_sim"]
    bb, mb = mc.get_amplitude_bounds([poly])
    cs = ctrl.all_moves("/tmp/golden", bb.bounds, mb.bounds, ctrl.SimpleControl)
    if per_cycle is not None:
        for c, k in zip(cs, per_cycle):
            c.move.num_per_cycle = k
    run = cs if order is None else [cs[i] for i in order]  # the controller list as the caller ordered it
    sh.c_srand(srand_seed)
    mcs.mc_sim([poly], df, steps, run, field, mu_adjust, np_seed)
    out = spec_to_npz(spec)
    out["steps"], out["srand_seed"], out["np_seed"], out["mu_adjust"] = steps, srand_seed, np_seed, mu_adjust
    out["per_cycle"] = np.array([c.move.num_per_cycle for c in cs])
    if order is not None:
        out["order"] = np.array(order)
    out["final_r"] = np.asarray(poly.r).copy()
    out["final_t3"] = np.asarray(poly.t3).copy()
    out["final_t2"] = np.asarray(poly.t2).copy()
    out["final_states"] = np.asarray(poly.states).copy()
    out["final_density"] = np.asarray(field.density).copy()
    out["num_attempt"] = np.array([c.move.num_attempt for c in cs])
    out["num_success"] = np.array([c.move.num_success for c in cs])
    out["amp_move"] = np.array([c.move.amp_move for c in cs])
    out["amp_bead"] = np.array([c.move.amp_bead for c in cs])
    out["acceptance_rate"] = np.array([c.move.acceptance_tracker.acceptance_rate for c in cs])
    out["E_field"] = field.compute_E(poly)
    out["E_poly"] = poly.compute_E()
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, "success", out["num_success"], "E", out["E_field"], out["E_poly"])


def golden_csv(name, spec, polymer_name):
    """CSV snapshot of the spec's polymer written by the reference's own writer
    (PolymerBase.to_csv, polymers.pyx:575-684)."""
    import warnings
    poly, _, _, _ = O.ref_objects(spec)
    poly.name = polymer_name
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        poly.to_csv(str(OUT / f"{name}.csv"))
    np.savez_compressed(OUT / f"{name}_spec.npz", **spec_to_npz(spec))


if __name__ == "__main__":
    only = sys.argv[1:]  # e.g. `make_golden.py csv`: just the snapshot CSVs
    if only == ["dc"]:
        # DetailedChromatin (polymers.pyx:2455-2607): nucleosomes with entry / exit points and an exit frame.
        # bead_rad is the nucleosome's (4.19 nm); linkers of 16.5 nm, 147 and 127 bp wrapped.
        dc = dict(O.make_spec(N=60, nb=1, seed=71, bead_rad=4.1899999999999995), lt=100.0, bp_wrap=147.0)
        dc2 = dict(O.make_spec(N=50, nb=2, seed=72, cross_talk=-1.0, bead_rad=4.1899999999999995), lt=60.0, bp_wrap=127.0)
        golden_static("static_dc", dc)
        golden_moves("moves_dc", dc, 200, 161)
        golden_moves("moves_dc2", dc2, 150, 162)
        golden_mc_sim("mcsim_dc", dict(O.make_spec(N=60, nb=1, seed=71, random_states=False, bead_rad=4.1899999999999995),
                                       lt=100.0, bp_wrap=147.0), 4, 27, 37)
        # DetailedChromatin2 (polymers.pyx:2627-2735): the same frames, bonds between the bead centres
        dc3 = dict(O.make_spec(N=60, nb=1, seed=73, bead_rad=4.1899999999999995), lt=100.0, bp_wrap=147.0, no_diameter=1)
        golden_moves("moves_dc3", dc3, 200, 163)
        golden_mc_sim("mcsim_dc3", dict(O.make_spec(N=60, nb=1, seed=73, random_states=False, bead_rad=4.1899999999999995),
                                        lt=100.0, bp_wrap=147.0, no_diameter=1), 4, 28, 38)
        sys.exit(0)
    if only == ["order"]:
        # a controller list that is not in all_moves' order (mc_sim.pyx:92 walks the list as given)
        golden_mc_sim("mcsim_order", O.make_spec(N=200, nb=1, seed=81, random_states=False), 5, 29, 39,
                      order=[4, 2, 0, 3, 1])
        sys.exit(0)
    if only == ["ff"]:
        # fast_field = 1 (fields.pyx:577-671, 1235-1368): positions quantised to n_points sub-bins per voxel edge
        ff = O.make_spec(N=300, nb=1, seed=51)
        ff["field"] = dict(ff["field"], fast_field=1, n_points=1000)
        ff2 = O.make_spec(N=200, nb=2, seed=52, cross_talk=-1.0, confine="", grid=6)
        ff2["field"] = dict(ff2["field"], fast_field=1, n_points=37)  # odd: rounded up to 38; periodic box
        golden_moves("moves_ff", ff, 300, 151)
        golden_moves("moves_ff2", ff2, 250, 152)
        ffs = O.make_spec(N=300, nb=1, seed=51, random_states=False)
        ffs["field"] = dict(ffs["field"], fast_field=1, n_points=1000)
        golden_mc_sim("mcsim_ff", ffs, 10, 26, 36)
        sys.exit(0)
    if only == ["av"]:
        # per-voxel accessible volumes (assume_fully_accessible = 0, fields.pyx:714-951): the voxels cut by
        # the confining sphere are smaller, every w / V_access of the polymer's outer shell changes
        av = O.make_spec(N=300, nb=1, seed=41)
        av["field"] = dict(av["field"], assume_fully_accessible=0)
        av2 = O.make_spec(N=200, nb=2, seed=42, cross_talk=-1.0, grid=7)
        av2["field"] = dict(av2["field"], assume_fully_accessible=0)
        golden_static("static_av", av)
        golden_static("static_av2", av2)
        golden_moves("moves_av", av, 300, 141)
        golden_moves("moves_av2", av2, 250, 142)
        avs = O.make_spec(N=300, nb=1, seed=41, random_states=False)
        avs["field"] = dict(avs["field"], assume_fully_accessible=0)
        golden_mc_sim("mcsim_av", avs, 10, 25, 35)
        sys.exit(0)
    # snapshot CSVs: two-binder chromatin and a null_reader SSWLC (lp != 53 -> the SSWLC class)
    golden_csv("snapshot_chromatin", O.make_spec(N=40, nb=2, seed=11, cross_talk=-1.5), "Chr-1")
    golden_csv("snapshot_sswlc", O.make_spec(N=25, nb=1, seed=12, binders=[dict(O.NULL_READER)], confine="",
                                             grid=4, random_states=False, lp=10.0), "homopolymer")
    if only == ["csv"]:
        sys.exit(0)
    # SSTWLC (twist, polymers.pyx:1889-2319): HP1 chain with a twist persistence length
    tw = dict(O.make_spec(N=220, nb=1, seed=7), lt=100.0)
    tw2 = dict(O.make_spec(N=180, nb=2, seed=8, cross_talk=-1.0, lp=30.0), lt=60.0)
    golden_static("static_tw", tw)
    golden_static("static_tw2", tw2)
    golden_moves("moves_tw", tw, 400, 105)
    golden_moves("moves_tw2", tw2, 300, 106)
    golden_mc_sim("mcsim_tw", dict(O.make_spec(N=220, nb=1, seed=7, random_states=False), lt=100.0), 12, 24, 34)
    if only == ["twist"]:
        sys.exit(0)
    # C1-like: homopolymer, null_reader, periodic box (confine_type="")
    c1 = O.make_spec(N=200, nb=1, seed=1, binders=[dict(O.NULL_READER)], confine="", grid=8,
                     random_states=False)
    # C2-like: HP1 chromatin in a spherical confinement
    c2 = O.make_spec(N=300, nb=1, seed=2)
    # C3-like: HP1 + PRC1 with cross-talk
    c3 = O.make_spec(N=250, nb=2, seed=3, cross_talk=-1.5)
    # over-dense box: exercises the vf_limit branch and confinement rejections
    c4 = O.make_spec(N=150, nb=1, seed=4, grid=4, vf_limit=0.02)
    golden_static("static_c1", c1)
    golden_static("static_c2", c2)
    golden_static("static_c3", c3)
    golden_static("static_c4", c4)
    golden_moves("moves_c1", c1, 300, 101)
    golden_moves("moves_c2", c2, 400, 102)
    golden_moves("moves_c3", c3, 400, 103)
    golden_moves("moves_c4", c4, 200, 104)
    golden_mc_sim("mcsim_c1", c1, 12, 21, 31)
    golden_mc_sim("mcsim_c2", O.make_spec(N=300, nb=1, seed=2, random_states=False), 15, 22, 32, mu_adjust=0.8)
    golden_mc_sim("mcsim_c3", c3, 10, 23, 33)
