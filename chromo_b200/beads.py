"""Bead objects (host-side).  Mirrors the parts of chromo/beads.py the hot path
reads: `beads[0].vol` (beads.py:142, 415; used by fields.pyx:1820) and
`beads[i].binders[b]` (beads.py:74-76: resolved BY NAME to the live binder
singletons; used by polymers.pyx:1487-1537).  Collision geometry (Prism,
DetailedNucleosome, GJK) is out of scope."""
import numpy as np

from .binders import get_by_name


class Bead:
    def __init__(self, id_, r, t3=None, t2=None, states=None, binder_names=None):
        self.id = id_
        self.r = r
        self.t3 = t3
        self.t2 = t2
        self.states = states
        self.binder_names = binder_names
        if binder_names is not None:
            self.binders = [get_by_name(name) for name in binder_names]
        else:
            self.binders = None


class GhostBead(Bead):
    """Bead of an SSWLC (beads.py:93-200)."""

    def __init__(self, id_, r, *, t3=None, t2=None, states=None, binder_names=None, rad=5, **kwargs):
        super().__init__(id_, r, t3, t2, states, binder_names)
        self.rad = rad
        self.vol = (4 / 3) * np.pi * rad ** 3
        self.kwargs = kwargs


class Nucleosome(Bead):
    """Bead of a Chromatin fiber (beads.py:369-470)."""

    def __init__(self, id_, r, *, t3, t2, states=None, binder_names=None, rad=5):
        super().__init__(id_, r, t3, t2, states, binder_names)
        self.rad = rad
        self.vol = (4 / 3) * np.pi * rad ** 3


class BeadMap:
    """`polymer.beads`: the reference builds one Python object per bead
    (polymers.pyx:1000-1012, 1874-1886); for 10^4-10^5 beads x 10^3 replicas we
    build them on demand instead -- same indexing, same attributes."""

    def __init__(self, poly, cls):
        self._poly, self._cls = poly, cls

    def __len__(self):
        return self._poly.num_beads

    def __getitem__(self, i):
        p = self._poly
        if i < 0 or i >= p.num_beads:
            raise KeyError(i)
        return self._cls(id_=i, r=p.r[i], t3=p.t3[i], t2=p.t2[i], states=p.states[i],
                         binder_names=p.binder_names, rad=p.bead_rad)

    def __iter__(self):
        return iter(range(len(self)))

    def keys(self):
        return range(len(self))

    def values(self):
        return (self[i] for i in range(len(self)))

    def items(self):
        return ((i, self[i]) for i in range(len(self)))
