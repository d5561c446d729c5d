"""The reference's histone-mark tracks (chromo/chemical_mods/*.txt), re-encoded.

The reference's simulation scripts read one 0/1/2 methylation count per
nucleosome bead from text files (`Chromatin.load_seqs`, polymers.pyx:1770-1812;
doc/source/one_mark_coarse.py:62-77) and give a polymer of N beads the first N
entries.  `chromo_b200/data/chemical_mods.npz` holds the same two tracks as
int8 arrays (tools/make_marks_fixture.py), so boxes without the reference tree
build the same chromatin."""
from pathlib import Path

import numpy as np

_PATH = Path(__file__).resolve().parents[1] / "data" / "chemical_mods.npz"
_CACHE = {}
TRACKS = ("H3K9me3", "H3K27me3")


def track(name: str) -> np.ndarray:
    """The whole track (451,692 beads) as int64, as np.loadtxt gives it to the reference."""
    if name not in _CACHE:
        with np.load(_PATH) as z:
            if name not in z.files:
                raise KeyError(f"no mark track {name!r}; have {TRACKS}")
            _CACHE[name] = z[name].astype(np.int64)
    return _CACHE[name]


def first_beads(num_beads: int, names=("H3K9me3",), replicas=None, start: int = 0) -> np.ndarray:
    """[N, len(names)] (or [R, N, len(names)]) marks: entries start .. start+N of each track, the same for
    every replica -- what `chemical_mods=...[:N]` does in the reference's scripts."""
    cols = []
    for n in names:
        t = track(n)
        if start + num_beads > len(t):
            raise ValueError(f"track {n} has {len(t)} beads; asked for {start}+{num_beads}")
        cols.append(t[start:start + num_beads])
    a = np.stack(cols, axis=1)
    if replicas is None:
        return a
    return np.ascontiguousarray(np.broadcast_to(a, (replicas,) + a.shape))
