"""Shared helpers for the test-suite (golden loading, tolerances)."""
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"

# north_star tolerance: energies / dE within 1e-9 relative in fp64.  The small
# absolute term covers sums that cancel to ~0 (summation ORDER over touched
# bins differs from the reference's Python-set order by construction).
RTOL = 1e-9
ATOL = 1e-11


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    spec = dict(meta)
    for k in ("r", "t3", "t2", "states", "mods", "bead_length"):
        spec[k] = z[k].copy()
    data = {k: z[k] for k in z.files}
    return spec, data


def close(a, b, rtol=RTOL, atol=ATOL):
    return abs(a - b) <= rtol * max(abs(a), abs(b)) + atol


def split(flat, counts):
    out, o = [], 0
    for c in counts:
        out.append(flat[o:o + c])
        o += c
    return out


def huge_scale(density, dtrial, touched, bead_vol, vf_limit):
    """Magnitude of the E_HUGE (1e99, fields.pyx:32) terms taking part in a
    field dE.  When a touched bin is above `vf_limit` the reference adds and
    subtracts 1e99*phi terms; their sum cancels to ~1e-16 of their magnitude
    and then depends on the summation order, so the comparison tolerance has
    to be relative to that magnitude, not to the cancelled result."""
    if len(touched) == 0:
        return 0.0
    vf0 = density[touched, 0] * bead_vol
    vf1 = vf0 + dtrial[touched, 0] * bead_vol
    lim = float(np.float32(vf_limit))
    big = np.concatenate([vf0[vf0 > lim], vf1[vf1 > lim]])
    return float(1e99 * np.abs(big).sum()) if len(big) else 0.0


def close_dE(a, b, scale=0.0, rtol=RTOL, atol=ATOL):
    return abs(a - b) <= rtol * max(abs(a), abs(b)) + atol + 1e-13 * scale
