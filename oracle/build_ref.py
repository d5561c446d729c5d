#!/usr/bin/env python
"""Build the UNMODIFIED reference (JosephWakim/chromo, Cython) into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (chromo_b200/) imports this.

What it does:
  1. copies /root/reference/chromo to a scratch dir under /tmp (the reference
     tree is read-only and must not be written to);
  2. applies the three one-line *toolchain compatibility* patches that the
     container's Cython 3.3 / pandas 3 need (no arithmetic is touched):
       - chromo/mc/moves.pxd:41-43 + moves.pyx:345  `cpdef list move_list`
         (Cython 3: "Variables cannot be declared with cpdef")
       - chromo/binders.pyx:237  DataFrame.append -> pd.concat
       - chromo/polymers.pyx:665  `df[name] = arr_temp` with an (N, 1) object
         array (pandas 3: "Buffer has wrong number of dimensions") ->
         `arr_temp[:, 0]`; only the CSV snapshot writer (to_dataframe) uses it
  3. adds `oracle_shim.pyx`, a thin `def` wrapper around the reference's
     `cdef` methods (compute_dE, propose, update_affected_densities) so tests
     can call them from Python;
  4. cythonizes the 8 extension modules of the reference's setup.py:14-23
     (language=c++, cdivision=False, language_level=2) with g++ -O2;
  5. installs the built package (compiled .so + the reference's .py files,
     minus the 3.5 MB chemical_mods data) into oracle/_ref/, plus an empty
     matplotlib stub (chromo/mc/__init__.py -> util/poly_stat.py imports it).

oracle/_ref/ is git-ignored (never part of the history) but NOT
gpurun-ignored, so the built reference travels to the GPU box.

Usage:  python oracle/build_ref.py [--force]
"""
import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_SRC = Path(os.environ.get("CHROMO_REFERENCE", "/root/reference"))
DEST = HERE / "_ref"
STAMP = DEST / ".built"

SHIM = '''\
# cython: language_level=3
"""def-wrappers around the reference's cdef hot-path methods (test shim)."""
cimport chromo.fields as fld
cimport chromo.polymers as ply
from chromo.mc.moves cimport MCAdapter
from libc.stdlib cimport rand, srand, RAND_MAX

def field_dE(fld.FieldBase f, ply.PolymerBase p, long[:] inds, long n,
             bint state_change):
    return f.compute_dE(p, inds, n, 20, state_change)

def poly_dE(ply.PolymerBase p, str name, long[:] inds, long n):
    return p.compute_dE(name, inds, n)

def propose(MCAdapter a, ply.PolymerBase p):
    return a.propose(p)

def commit(fld.FieldBase f):
    f.update_affected_densities()

def c_srand(unsigned int seed):
    srand(seed)

def c_rand():
    return rand()
'''

SETUP = '''\
import numpy as np
from setuptools import setup, Extension
from Cython.Build import cythonize
paths = ["chromo/mc/move_funcs.pyx", "chromo/polymers.pyx", "chromo/fields.pyx",
         "chromo/binders.pyx", "chromo/util/bead_selection.pyx",
         "chromo/util/linalg.pyx", "chromo/mc/mc_sim.pyx", "chromo/mc/moves.pyx",
         "oracle_shim.pyx"]
exts = [Extension(p.split(".")[0].replace("/", "."), sources=[p], language="c++",
                  extra_compile_args=["-O2", "-w"]) for p in paths]
setup(name="chromo_ref", ext_modules=cythonize(
    exts, nthreads=%d,
    compiler_directives={"cdivision": False, "language_level": 2}),
    include_dirs=[np.get_include(), "chromo", "chromo/util", "chromo/mc"],
    script_args=["build_ext", "--inplace", "-j", "%d"])
'''


def _patch(path: Path, old: str, new: str, count: int = 1):
    text = path.read_text()
    if old not in text:
        raise RuntimeError(f"patch anchor not found in {path}: {old!r}")
    path.write_text(text.replace(old, new, count))


def build(force: bool = False) -> bool:
    """Return True when oracle/_ref is usable (built now or earlier)."""
    if STAMP.exists() and not force:
        return True
    if not (REF_SRC / "chromo" / "fields.pyx").exists():
        return False  # e.g. on the GPU box: use whatever was prebuilt
    ncpu = max(1, os.cpu_count() or 1)
    tmp = Path(tempfile.mkdtemp(prefix="chromo_ref_build_"))
    try:
        work = tmp / "src"
        shutil.copytree(REF_SRC / "chromo", work / "chromo",
                        ignore=shutil.ignore_patterns("chemical_mods", "__pycache__"))
        for root, dirs, files in os.walk(work):
            for n in dirs + files:
                os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
        # patch 1: Cython 3 rejects `cpdef list move_list = [...]`
        pxd = work / "chromo/mc/moves.pxd"
        _patch(pxd, "cpdef list move_list = [\n    crank_shaft, end_pivot, slide,"
                    " tangent_rotation, change_binding_state\n]\n", "")
        _patch(work / "chromo/mc/moves.pyx", "cpdef list move_list = [", "move_list = [")
        # patch 2: pandas >= 2 removed DataFrame.append
        _patch(work / "chromo/binders.pyx",
               "df = df.append(binder.dict(), ignore_index=True)",
               "df = pd.concat([df, pd.DataFrame([binder.dict()])], ignore_index=True)")
        # patch 3: pandas 3 refuses an (N, 1) object array as a column (snapshot writer only)
        _patch(work / "chromo/polymers.pyx", "            df[name] = arr_temp\n",
               "            df[name] = arr_temp[:, 0]\n")
        (work / "oracle_shim.pyx").write_text(SHIM)
        (work / "setup_ref.py").write_text(SETUP % (ncpu, ncpu))
        env = dict(os.environ)
        env.pop("PYTHONPATH", None)
        subprocess.run([sys.executable, "setup_ref.py"], cwd=work, check=True, env=env,
                       stdout=subprocess.DEVNULL)
        # install: package + shim .so into oracle/_ref
        if DEST.exists():
            shutil.rmtree(DEST)
        DEST.mkdir(parents=True)
        shutil.copytree(work / "chromo", DEST / "chromo",
                        ignore=shutil.ignore_patterns("*.cpp", "*.html", "*.c", "__pycache__"))
        for so in work.glob("oracle_shim*.so"):
            shutil.copy2(so, DEST / so.name)
        stub = DEST / "_stubs" / "matplotlib"
        stub.mkdir(parents=True)
        (stub / "__init__.py").write_text("")
        (stub / "pyplot.py").write_text("")
        STAMP.write_text("ok\n")
        return True
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def available() -> bool:
    return STAMP.exists()


def activate():
    """Put oracle/_ref on sys.path (after which `import chromo` is the reference)."""
    if not available():
        raise ImportError("oracle/_ref not built; run python oracle/build_ref.py")
    for p in (str(DEST / "_stubs"), str(DEST)):
        if p not in sys.path:
            sys.path.insert(0, p)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "built" if ok else "reference sources not present")
    sys.exit(0 if ok else 1)
