"""The known-answer tests the REFERENCE's own test-suite holds for this path, restated
against the CPU oracle: rotation / translation matrices (tests/test_linalg.py:12-65), the deterministic
part of every move (tests/test_moves.py:16-282), the voxel super-index convention of the 5x5x5 neighbour
table (tests/test_fields.py:62-96) and the bead-selection distributions
(tests/test_bead_selection.py:213-313).  The expected values are the reference tests' own literals.

These pin the oracle; the CUDA path is compared with the oracle on the same quantities in
tests/test_parity.py (trial rows of every proposed move, bin indices and weights, selected indices)."""
import ctypes as C

import numpy as np
import pytest

import oracle as O

_pd, _pl = O._pd, O._pl


def rotation(axis, point, angle):
    m = np.zeros(16)
    O.lib().oc_rotation_matrix(O._p(np.ascontiguousarray(axis, dtype=float), _pd),
                               O._p(np.ascontiguousarray(point, dtype=float), _pd), float(angle), O._p(m, _pd))
    return m.reshape(4, 4)


def transformed(r, t3, t2, mat, inds):
    """transform_r_t3_t2 (move_funcs.pyx:121-154) through the oracle: trial rows of a chain."""
    N = len(r)
    spec = O.make_spec(N=max(N, 4), nb=1, seed=0, grid=4, confine="")
    s = O.OracleSim(spec)
    for dst, src in ((s.r, r), (s.t3, t3), (s.t2, t2)):
        dst[:N] = src
    for dst, src in ((s.r_trial, r), (s.t3_trial, t3), (s.t2_trial, t2)):
        dst[:N] = src
    inds = np.ascontiguousarray(inds, dtype=np.int64)
    O.lib().oc_transform_rows(C.byref(s.s), O._p(np.ascontiguousarray(mat, dtype=float).ravel(), _pd),
                              O._p(inds, _pl), len(inds))
    return s.r_trial[:N].copy(), s.t3_trial[:N].copy(), s.t2_trial[:N].copy()


# ---- tests/test_linalg.py ---------------------------------------------------------------------
def test_z_axis_rotation():
    ang = np.pi / 4
    m = rotation([0, 0, 1.0], [0, 0, 0.0], ang)
    inv_inv = np.linalg.inv(rotation([0, 0, 1.0], [0, 0, 0.0], -ang))
    want = np.identity(4)
    want[0:2, 0:2] = [[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]
    assert np.allclose(m, inv_inv) and np.allclose(m, want)


def test_translation():
    m = np.identity(4)
    m[:3, 3] = [0.1, 0.2, 0.3]  # generate_translation_mat linalg.pyx:172-199
    r, _, _ = transformed(np.ones((2, 3)), np.zeros((2, 3)), np.zeros((2, 3)), m, [0, 1])
    assert np.allclose(r[0], [1.1, 1.2, 1.3])


# ---- tests/test_moves.py ------------------------------------------------------------------------
c = np.sqrt(0.5)
R5 = np.array([[1, 0, 0], [2, 0, 0], [3, 0, 0], [3, -1, 0], [3, -2, 0]], dtype=float)


def test_deterministic_end_pivot():
    t3 = np.array([[1, 0, 0], [1, 0, 0], [c, -c, 0], [0, -1, 0], [0, -1, 0]], dtype=float)
    t2 = np.tile([0, 0, 1.0], (5, 1))
    m = rotation([1.0, 0, 0], R5[0], -np.pi / 2)
    r, a, b = transformed(R5, t3, t2, m, np.arange(5))
    assert np.allclose(r, [[1, 0, 0], [2, 0, 0], [3, 0, 0], [3, 0, 1], [3, 0, 2]])
    assert np.allclose(a, [[1, 0, 0], [1, 0, 0], [c, 0, c], [0, 0, 1], [0, 0, 1]])
    assert np.allclose(b, np.tile([0, 1.0, 0], (5, 1)))


def test_deterministic_slide_move():
    m = np.identity(4)
    m[:3, 3] = [1, 2, 3.5]
    r, _, _ = transformed(R5, np.zeros((5, 3)), np.zeros((5, 3)), m, np.arange(5))
    assert np.allclose(r, [[2, 2, 3.5], [3, 2, 3.5], [4, 2, 3.5], [4, 1, 3.5], [4, 0, 3.5]])


def test_deterministic_tangent_rotation():
    m = rotation(np.array([1, 1, 1.0]) / np.sqrt(3), [1, 2, 3.0], 2 * np.pi / 3)
    r, a, b = transformed(np.array([[1, 2, 3.0], [9, 9, 9.0]]), np.array([[0, 0, 1.0], [1, 0, 0]]),
                          np.array([[0, 1.0, 0], [0, 1, 0]]), m, [0])
    assert np.allclose(r[0], [1, 2, 3]) and np.allclose(a[0], [1, 0, 0]) and np.allclose(b[0], [0, 0, 1])
    assert np.allclose(r[1], [9, 9, 9])  # rows outside `inds` are untouched


def test_deterministic_crank_shaft_move():
    r = np.array([[1, 0, 0], [2, 0, 0], [3, 0, 0], [3, -1, 0], [3, -2, 0], [4, -2, 0], [5, -2, 0], [5, -1, 0],
                  [5, 0, 0], [6, 0, 0], [7, 0, 0]], dtype=float)
    t3 = np.array([[1, 0, 0], [1, 0, 0], [c, -c, 0], [0, -1, 0], [1, 0, 0], [1, 0, 0], [-c, c, 0], [0, 1, 0],
                   [1, 0, 0], [1, 0, 0], [1, 0, 0]], dtype=float)
    axis = r[2] - r[8]
    m = rotation(axis / np.linalg.norm(axis), r[0], np.pi / 2)
    rt, a, _ = transformed(r, t3, np.zeros_like(r), m, np.arange(3, 8))
    want_r = r.copy()
    want_r[3:8] = [[3, 0, 1], [3, 0, 2], [4, 0, 2], [5, 0, 2], [5, 0, 1]]
    want_t = t3.copy()
    want_t[3:8] = [[0, 0, 1], [1, 0, 0], [1, 0, 0], [-c, 0, -c], [0, 0, -1]]
    assert np.allclose(rt, want_r) and np.allclose(a, want_t)


# ---- tests/test_fields.py: super-index convention ix + nx*iy + nx*ny*iz with periodic wrap ----------
def test_voxel_super_indices_on_a_5x5x5_grid():
    spec = O.make_spec(N=4, nb=1, seed=0, grid=5, confine="")
    s = O.OracleSim(spec)
    W = spec["field"]["x_width"]
    d = W / 5
    # the 8 voxels a bead contributes to are neighbours (or the voxel itself) in the reference's table
    nbrs_0 = {24, 20, 21, 4, 0, 1, 9, 5, 6, 29, 25, 26, 34, 30, 31, 49, 45, 46, 104, 100, 101, 109, 105, 106,
              120, 121, 124}
    nbrs_12 = {6, 7, 8, 11, 12, 13, 16, 17, 18, 106, 107, 108, 111, 112, 113, 116, 117, 118, 31, 32, 33, 36, 37,
               38, 41, 42, 43}
    # a bead just inside the lower corner of voxel 0: its lower neighbours wrap to index 4 on every axis
    corner = -W / 2 + 0.25 * d
    idx, w = s.bin_point([corner, corner, corner])
    assert set(idx) == {0, 4, 20, 24, 100, 104, 120, 124} and set(idx) <= nbrs_0
    assert np.isclose(w.sum(), 1.0)
    # a bead in voxel 12 = (2, 2, 0), low in z: wraps to the top layer (100 + ...)
    idx, _ = s.bin_point([-W / 2 + 2.75 * d, -W / 2 + 2.25 * d, -W / 2 + 0.25 * d])
    assert set(idx) == {12, 13, 7, 8, 112, 113, 107, 108} and set(idx) <= nbrs_12
    # x is the fastest index (bit 0 of the corner number), then y, then z
    idx, _ = s.bin_point([-W / 2 + 1.75 * d, -W / 2 + 3.75 * d, -W / 2 + 2.75 * d])
    assert list(idx) == [1 + 5 * 3 + 25 * 2 + dx + 5 * dy + 25 * dz for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)]


# ---- tests/test_bead_selection.py: exponential windows -------------------------------------------------
def _draws(fn, n, *args):
    g = O.GlibcRand()
    O.lib().oc_srand(C.byref(g), 12345)
    return np.array([fn(C.byref(g), *args) for _ in range(n)])


def test_from_left_and_from_right():
    N = 10000
    left = _draws(O.lib().oc_from_left, 20000, N, N)
    right = _draws(O.lib().oc_from_right, 20000, N, N)
    for sel in (left, right):
        assert sel.min() >= 0 and sel.max() <= N
    hl, _ = np.histogram(left, bins=20, range=(0, N))
    hr, _ = np.histogram(right, bins=20, range=(0, N))
    assert hl[0] > hl[-1] and hr[0] < hr[-1]            # the reference's y[0] > y[-1] / y[0] < y[-1]
    assert np.all(np.diff(hl[:10]) <= 0.05 * hl[0])      # decaying from the chosen end
    # -log10(u) * window * 0.45: the mean distance from the end is ~0.195 N (bead_selection.pyx:58-60)
    assert abs(left.mean() / N - 0.195) < 0.02 and abs((N - right.mean()) / N - 0.195) < 0.02


def test_from_point():
    N = 10000
    for ind0 in (0, 17, 5000, 9999):
        sel = _draws(O.lib().oc_from_point, 5000, N, N, ind0)
        assert sel.min() >= 0 and sel.max() <= N
        assert np.median(np.abs(sel - ind0)) < 0.2 * N    # concentrated around the chosen bead
        if 100 < ind0 < N - 100:
            frac_right = np.mean(sel > ind0)
            assert 0.4 < frac_right < 0.6                 # either side with probability 1/2


# ---- the same convention through the CUDA path (full density recompute of beads placed at the KAT points) ----
def test_voxel_super_indices_through_the_kernels(backend):
    from gpu_common import engine_from_spec
    spec = O.make_spec(N=4, nb=1, seed=0, grid=5, confine="")
    W = spec["field"]["x_width"]
    d = W / 5
    s = O.OracleSim(spec)
    cases = [([-W / 2 + 0.25 * d] * 3, {0, 4, 20, 24, 100, 104, 120, 124}),
             ([-W / 2 + 2.75 * d, -W / 2 + 2.25 * d, -W / 2 + 0.25 * d], {12, 13, 7, 8, 112, 113, 107, 108}),
             ([-W / 2 + 1.75 * d, -W / 2 + 3.75 * d, -W / 2 + 2.75 * d], {66, 67, 71, 72, 91, 92, 96, 97})]
    for point, want in cases:
        spec["r"] = np.tile(np.asarray(point, dtype=float), (4, 1))
        e = engine_from_spec(spec, R=1)
        dens = e.density()[0]
        assert set(np.nonzero(dens[:, 0])[0]) == want
        idx, w = s.bin_point(point)
        vol_bin = d ** 3
        assert np.allclose(dens[idx, 0] * vol_bin, 4 * w, rtol=1e-12)  # four beads, the oracle's weights
        e.close()
