#!/usr/bin/env python
"""Re-encode the reference's two histone-mark tracks (chromo/chemical_mods/HNCFF683HCZ_H3K9me3_methyl.txt,
ENCFF919DOR_H3K27me3_methyl.txt: one 0/1/2 count per 200-bp nucleosome bead) as one compressed
int8 .npz under chromo_b200/data/, so that the bench and the tests build their chromatin from the SAME marks
the reference's own scripts load (one_mark_coarse.py, two_mark_factorial.py) on boxes where /root/reference
does not exist.  Data only -- no reference source is copied.

    python tools/make_marks_fixture.py [/root/reference]
"""
import sys
from pathlib import Path

import numpy as np

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference") / "chromo" / "chemical_mods"
out = Path(__file__).resolve().parents[1] / "chromo_b200" / "data" / "chemical_mods.npz"
tracks = {"H3K9me3": "HNCFF683HCZ_H3K9me3_methyl.txt", "H3K27me3": "ENCFF919DOR_H3K27me3_methyl.txt"}
arrs = {}
for name, fn in tracks.items():
    a = np.loadtxt(ref / fn, dtype=np.int64)
    assert a.min() >= 0 and a.max() <= 2
    arrs[name] = a.astype(np.int8)
    print(name, fn, len(a), np.bincount(a) / len(a))
np.savez_compressed(out, **arrs, files=np.array([tracks[k] for k in arrs]))
print(out, out.stat().st_size, "bytes")
