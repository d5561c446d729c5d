#!/usr/bin/env python
"""Golden vectors for the coarse-grain / refine pipeline, produced by running the
REFERENCE ITSELF (chromo/util/rediscretize.py from oracle/_ref, the unmodified
build of /root/reference made by oracle/build_ref.py).

Run in the authoring container only:  python tests/golden/make_golden_rediscretize.py
Output: tests/golden/rediscretize.npz (inputs, the Gaussian deviates numpy's global
generator handed the reference, and the reference's outputs)."""
import contextlib
import io
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT / "oracle"))
import build_ref  # noqa: E402

build_ref.activate()
import chromo.util.rediscretize as ref  # noqa: E402
import chromo.polymers as ply  # noqa: E402
import chromo.binders as bnd  # noqa: E402
import chromo.fields as fld  # noqa: E402

sys.path.insert(0, str(ROOT / "oracle"))
import rediscretize_oracle as RO  # noqa: E402

out = {}
rng = np.random.default_rng(2024)


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


# ---- coarse-graining: the four building blocks --------------------------------------
cg_cases = [(103, 5, 2), (64, 8, 1), (40, 1, 1), (57, 57, 3), (1000, 12, 2)]
for i, (N, k, nb) in enumerate(cg_cases):
    r = np.cumsum(rng.standard_normal((N, 3)), axis=0) * 16.5
    t3 = rng.standard_normal((N, 3))
    t3 /= np.linalg.norm(t3, axis=1)[:, None]
    if i == 1:
        t3[8:16] = [1.0, 0.0, 0.0]   # an interval whose mean tangent is exactly e_x
        t3[16:24] = [0.0, -1.0, 0.0]
    st = rng.integers(0, 3, (N, nb))
    md = rng.integers(0, 3, (N, nb))
    iv = ref.get_cg_bead_intervals(N, k)
    t3_cg, t2_cg = ref.get_orientations_in_intervals(t3, iv)
    out.update({f"cg{i}_shape": np.array([N, k, nb]), f"cg{i}_r": r, f"cg{i}_t3": t3, f"cg{i}_states": st,
                f"cg{i}_mods": md, f"cg{i}_avg": ref.get_avg_in_intervals(r, iv), f"cg{i}_t3cg": t3_cg,
                f"cg{i}_t2cg": t2_cg, f"cg{i}_maj_states": ref.get_majority_state_in_interval(st, iv),
                f"cg{i}_maj_mods": ref.get_majority_state_in_interval(md, iv)})
out["cg_n"] = np.array(len(cg_cases))

# ---- get_cg_chromatin / get_cg_udf on real objects --------------------------------------
N, k = 600, 5
r = np.cumsum(rng.standard_normal((N, 3)), axis=0) * 6.0
r -= r.mean(axis=0)
t3 = rng.standard_normal((N, 3))
t3 /= np.linalg.norm(t3, axis=1)[:, None]
t2 = np.cross(t3, [1.0, 0, 0])
t2 /= np.linalg.norm(t2, axis=1)[:, None]
st = rng.integers(0, 3, (N, 1))
md = rng.integers(0, 3, (N, 1))
hp1 = bnd.get_by_name("HP1")
df = bnd.make_binder_collection([hp1])
poly = ply.Chromatin("c", r.copy(), bead_length=np.ones(N - 1) * 16.5, t3=t3.copy(), t2=t2.copy(), states=st.copy(),
                     binder_names=np.array(["HP1"]), chemical_mods=md.copy(),
                     chemical_mod_names=np.array(["H3K9me3"]))
R0 = float(np.max(np.linalg.norm(r, axis=1))) * 1.05
udf = fld.UniformDensityField([poly], df, 2.4 * R0, 12, 2.4 * R0, 12, 2.4 * R0, 12, confine_type="Spherical",
                              confine_length=R0, chi=1.0)
pcg = ref.get_cg_chromatin(poly, k)
ucg = ref.get_cg_udf(udf.dict_, df, k, [pcg])
out.update(dict(obj_r=r, obj_t3=t3, obj_t2=t2, obj_states=st, obj_mods=md, obj_k=np.array(k), obj_R0=np.array(R0),
                obj_cg_r=np.asarray(pcg.r), obj_cg_t3=np.asarray(pcg.t3), obj_cg_t2=np.asarray(pcg.t2),
                obj_cg_states=np.asarray(pcg.states), obj_cg_mods=np.asarray(pcg.chemical_mods),
                obj_cg_bead_length=np.asarray(pcg.bead_length),
                obj_cg_grid=np.array([ucg.nx, ucg.ny, ucg.nz, ucg.x_width, ucg.y_width, ucg.z_width,
                                      ucg.confine_length]),
                obj_cg_density=np.asarray(ucg.density)))

# ---- refined paths ---------------------------------------------------------------------------
rf_cases = [(6, 100, 2.0), (6, 100, 30.0), (4, 24, 2.0), (9, 200, 16.5), (3, 40, 5.0), (5, 17, 1.0), (12, 1000, 4.0)]
for i, (M, n_ref, sp) in enumerate(rf_cases):
    cg = np.cumsum(rng.standard_normal((M, 3)), axis=0) * 20.0
    L = RO.refine_layout(M, n_ref)
    np.random.seed(100 + i)
    path = quiet(ref.get_refined_path, cg, n_ref, sp)
    assert len(path) == L["points"], (len(path), L)
    state_after = np.random.get_state()[2]
    np.random.seed(100 + i)
    xi = np.random.standard_normal((L["draws"], 3))
    assert np.random.get_state()[2] == state_after, "draw count differs from the reference's"
    t = rng.standard_normal((M, 3))
    t /= np.linalg.norm(t, axis=1)[:, None]
    np.random.seed(200 + i)
    o3, o2 = quiet(ref.get_refined_orientations, t, n_ref)
    np.random.seed(200 + i)
    xo = np.random.standard_normal((L["draws"], 3))
    out.update({f"rf{i}_shape": np.array([M, n_ref]), f"rf{i}_spacing": np.array(sp), f"rf{i}_cg": cg,
                f"rf{i}_xi": xi, f"rf{i}_path": path, f"rf{i}_t": t, f"rf{i}_xo": xo, f"rf{i}_t3": o3,
                f"rf{i}_t2": o2})
out["rf_n"] = np.array(len(rf_cases))

# ---- spherical confinement ------------------------------------------------------------------------
for i, (N, rad) in enumerate([(300, 40.0), (50, 3.0), (7, 1.0), (2000, 150.0)]):
    r = np.cumsum(rng.standard_normal((N, 3)), axis=0) * 5.0
    out[f"cf{i}_r"] = r
    out[f"cf{i}_rad"] = np.array(rad)
    out[f"cf{i}_out"] = ref.enforce_spherical_confinement(r.copy(), rad)
out["cf_n"] = np.array(4)

# ---- refine_chromatin on the coarse-grained objects (no binding equilibration: geometry only) ----
np.random.seed(77)
n_ref = 601   # not a multiple of len(pcg.r) - 1, so the path has n_ref rows (see chromo_refined_num_points)
pref, uref = quiet(ref.refine_chromatin, pcg, n_ref, 16.5, np.ascontiguousarray(np.vstack([md, md[:1]])), ucg)
out.update(dict(rc_nref=np.array(n_ref), rc_mods=np.vstack([md, md[:1]]), rc_r=np.asarray(pref.r),
                rc_t3=np.asarray(pref.t3), rc_t2=np.asarray(pref.t2),
                rc_grid=np.array([uref.nx, uref.ny, uref.nz, uref.x_width, uref.y_width, uref.z_width,
                                  uref.confine_length])))

# ---- the two path primitives on their own (appended last: earlier vectors keep their random inputs) ----
for i, (N, tgt) in enumerate([(1, 2.0), (2, 0.5), (5, 8.0), (33, 3.0), (200, 16.5)]):
    p0, p1 = rng.standard_normal(3) * 10, rng.standard_normal(3) * 10
    np.random.seed(300 + i)
    out[f"bb{i}_out"] = quiet(ref.brownian_bridge, N, p0, p1, tgt)
    np.random.seed(400 + i)
    out[f"gw{i}_out"] = ref.gaussian_walk_from_point(p0, N, np.array([tgt] * N))
    out[f"bb{i}_in"] = np.concatenate([[N, tgt], p0, p1])
out["bb_n"] = np.array(5)

np.savez_compressed(Path(__file__).resolve().parent / "rediscretize.npz", **out)
print("wrote rediscretize.npz with", len(out), "arrays")
