"""Initial configurations, vectorised over replicas.

Distribution-equivalent to chromo/util/poly_paths.py (gaussian_walk 269-296,
confined_gaussian_walk 298-335 -- whose np.vstack loop is O(N^2) --, and
estimate_tangents_from_coordinates 506-545) but written for [R, N, 3] batches."""
import numpy as np


def in_confinement(points, confine_type, confine_length):
    """poly_paths.py:433-503 on an array of points."""
    if confine_type == "":
        return np.ones(points.shape[:-1], dtype=bool)
    if confine_type == "Spherical":
        return np.linalg.norm(points, axis=-1) <= confine_length
    if confine_type == "Cubical":
        return np.all(np.abs(points) <= confine_length / 2, axis=-1)
    raise ValueError("Confinement type " + confine_type + " not found.")


def gaussian_walk(num_steps, step_sizes, rng=None, replicas=None):
    rng = np.random.default_rng() if rng is None else rng
    shape = (num_steps, 3) if replicas is None else (replicas, num_steps, 3)
    steps = rng.standard_normal(shape)
    steps /= np.linalg.norm(steps, axis=-1, keepdims=True)
    pos = np.cumsum(steps * np.asarray(step_sizes)[..., :, None], axis=-2)
    zero = np.zeros(shape[:-2] + (1, 3))
    return np.concatenate([zero, pos], axis=-2)


def confined_gaussian_walk(num_points, step_sizes, confine_type, confine_length, rng=None, replicas=None):
    """Unit Gaussian-direction steps of the given lengths, each re-drawn while
    it would leave the confinement; all replicas advance together."""
    rng = np.random.default_rng() if rng is None else rng
    R = 1 if replicas is None else replicas
    step_sizes = np.broadcast_to(np.asarray(step_sizes, dtype=float), (num_points - 1,))
    pts = np.zeros((R, num_points, 3))
    for i in range(num_points - 1):
        todo = np.arange(R)
        while len(todo):
            step = rng.standard_normal((len(todo), 3))
            step *= step_sizes[i] / np.linalg.norm(step, axis=1, keepdims=True)
            cand = pts[todo, i] + step
            ok = in_confinement(cand, confine_type, confine_length)
            pts[todo[ok], i + 1] = cand[ok]
            todo = todo[~ok]
    return pts[0] if replicas is None else pts


def estimate_tangents_from_coordinates(coordinates):
    """t3 = normalised central differences; t2 = t3 x e_x (or t3 x e_y when
    t3 is parallel to e_x), normalised (poly_paths.py:506-545).  Accepts
    [N,3] or [R,N,3]."""
    c = np.asarray(coordinates, dtype=float)
    t3 = np.empty_like(c)
    t3[..., 1:-1, :] = c[..., 2:, :] - c[..., :-2, :]
    t3[..., 0, :] = c[..., 1, :] - c[..., 0, :]
    t3[..., -1, :] = c[..., -1, :] - c[..., -2, :]
    t3 /= np.linalg.norm(t3, axis=-1, keepdims=True)
    t2 = np.cross(t3, np.array([1.0, 0.0, 0.0]))
    bad = np.all(t2 == 0, axis=-1)
    if np.any(bad):
        t2[bad] = np.cross(t3[bad], np.array([0.0, 1.0, 0.0]))
    t2 /= np.linalg.norm(t2, axis=-1, keepdims=True)
    return t3, t2


def synthetic_marks(num_beads, num_marks, rng=None, p=(0.457, 0.084, 0.459), domain=40, replicas=None):
    """Blocky 0/1/2 methylation tracks with the marginal distribution of the
    reference's H3K9me3 track (chromo/chemical_mods/HNCFF683HCZ_H3K9me3_methyl.txt:
    45.7 / 8.4 / 45.9 % of 0 / 1 / 2) and ~`domain`-bead domains."""
    rng = np.random.default_rng() if rng is None else rng
    R = 1 if replicas is None else replicas
    out = np.zeros((R, num_beads, num_marks), dtype=np.int64)
    for r in range(R):
        for b in range(num_marks):
            nseg = int(2.5 * num_beads / domain) + 4
            lens = 1 + rng.geometric(1.0 / domain, size=nseg)
            vals = rng.choice(3, size=nseg, p=p)
            out[r, :, b] = np.repeat(vals, lens)[:num_beads]
    return out[0] if replicas is None else out
