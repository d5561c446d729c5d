#!/usr/bin/env python
"""Device time and achieved HBM bandwidth of the coarse-grain / refine kernels
(csrc/rediscretize.cu) at BASELINE.json's sizes.  Kernel time = CUDA events around
the launch inside the C-ABI call (`kernel_ms`); bytes = algorithmic bytes
(DESIGN.md 4.4).  Usage: python tools/rediscretize_bench.py [--out file.json]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import chromo_b200.util.rediscretize as rd  # noqa: E402


def peak():
    try:
        p = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        for k in ("hbm_gbs", "hbm_gbps"):
            if k in p:
                return float(p[k]), "measured:" + k
        for k, v in p.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v), "measured:" + k
    except Exception:
        pass
    return 6545.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="", help="run only the cases whose name starts with this (e.g. C2)")
    a = ap.parse_args()
    pk, src = peak()
    rng = np.random.default_rng(0)
    res = {"peak_gbps": pk, "peak_source": src, "cases": []}
    for name, R, N, k, nb in (("C2 1024 x 10000 beads, cg 5", 1024, 10000, 5, 1),
                              ("C3 1024 x 10000 beads, 2 binders, cg 10", 1024, 10000, 10, 2),
                              ("C4 16 x 400000 beads, cg 40", 16, 400000, 40, 1)):
        if a.only and not name.startswith(a.only):
            continue
        r = rng.standard_normal((R, N, 3))
        t3 = rng.standard_normal((R, N, 3))
        st = rng.integers(0, 3, (R, N, nb))
        M = N // k + (1 if N % k else 0)
        ms = []
        t0 = time.perf_counter()
        for _ in range(a.reps):
            ms.append(rd.coarse_grain_ensemble(r, t3, st, st, k)["kernel_ms"])
        wall = (time.perf_counter() - t0) / a.reps
        best = float(np.median(ms[1:])) if len(ms) > 1 else ms[0]
        byt = R * N * (48 + 16 * nb) + R * M * (72 + 16 * nb)
        res["cases"].append(dict(kernel="cg_reduce_kernel", case=name, kernel_ms=best, algorithmic_bytes=byt,
                                 gbps=byt / best / 1e6, frac=byt / best / 1e6 / pk, call_wall_ms=1e3 * wall,
                                 beads_per_s=R * N / (best * 1e-3)))
        # refine back to N + 1 beads from the coarse path (device-side deviates)
        cg = np.ascontiguousarray(rng.standard_normal((R, M, 3)).cumsum(axis=1) * 30)
        ms = []
        for _ in range(a.reps):
            _, _, t = rd._refine(cg, N + 1, 16.5, orientations=False, out_scale=1.0, seed=1)
            ms.append(t)
        best = float(np.median(ms[1:])) if len(ms) > 1 else ms[0]
        byt = R * (N + 1) * 24 + R * M * 24
        res["cases"].append(dict(kernel="refine_path_kernel", case=name + " -> refined", kernel_ms=best,
                                 algorithmic_bytes=byt, gbps=byt / best / 1e6, frac=byt / best / 1e6 / pk,
                                 beads_per_s=R * (N + 1) / (best * 1e-3)))
        x = rng.standard_normal((R, N, 3)) * 100
        t = np.zeros(1)
        ms = []
        for _ in range(a.reps):
            y = x.copy()
            from chromo_b200 import _lib
            _lib.check(_lib.lib().chromo_enforce_spherical_confinement(0, R, N, _lib.dptr(y), 150.0, _lib.dptr(t)))
            ms.append(float(t[0]))
        best = float(np.median(ms[1:])) if len(ms) > 1 else ms[0]
        byt = R * N * 48
        res["cases"].append(dict(kernel="confine_kernel", case=name, kernel_ms=best, algorithmic_bytes=byt,
                                 gbps=byt / best / 1e6, frac=byt / best / 1e6 / pk))
    s = json.dumps(res, indent=1)
    print(s)
    if a.out:
        Path(a.out).write_text(s)


if __name__ == "__main__":
    main()
