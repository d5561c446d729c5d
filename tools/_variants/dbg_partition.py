import sys, ctypes as C
sys.path[:0]=['/root/repo','/root/repo/oracle','/root/repo/tests']
import numpy as np, oracle as O
import os
from chromo_b200 import _lib
if os.environ.get('CB_LIB'): _lib.use_library(os.environ['CB_LIB'])
from gpu_common import engine_from_spec
spec = O.make_spec(N=300, nb=2, seed=5, grid=24, cross_talk=-0.7)
for cap in (128, 256):
    e = engine_from_spec(spec, R=1)
    e.set_table_capacity(cap)
    o = O.OracleSim(spec)
    e.srand(3), o.srand(3), e.numpy_seed(3), o.np_seed(3)
    mvs = O.make_moves(spec["N"], 16.5)
    rng = np.random.default_rng(0)
    bad = 0
    for it in range(40):
        m = int(rng.integers(0, 5))
        amp_bead, amp_move = int(rng.integers(40, 150)), 0.3 * (1 + m)
        inds = o.propose(m, amp_move, amp_bead)
        dEp = o.poly_dE(m, inds)
        dEf, touched = (0.0, np.zeros(0, dtype=np.int64)) if m == 3 else o.field_dE(inds, m == 4)
        with np.errstate(over="ignore"):
            acc = int(rng.uniform() < np.exp(-(dEp + dEf)))
        out = e.mc_step(0, m, amp_move, amp_bead, 1.0, 1, 0, acc)
        if m != 3:
            a, b = np.sort(out["touched"]), np.sort(touched)
            if not np.array_equal(a, b):
                bad += 1
                print("cap", cap, "it", it, "move", m, "n", len(inds), "passes", out["passes"], "U gpu/ora", len(a), len(b),
                      "missing", sorted(set(b) - set(a))[:8], "extra", sorted(set(a) - set(b))[:8], "dups", len(a) - len(set(a)))
        ip = inds.ctypes.data_as(O._pl)
        if acc:
            O.lib().oc_accept(C.byref(o.s), C.byref(mvs[m]), m, ip, len(inds))
            if m != 3:
                o.commit_field()
        else:
            O.lib().oc_reject(C.byref(o.s), C.byref(mvs[m]), m, ip, len(inds))
    print("cap", cap, "bad", bad)
    e.close()
