"""Polymer state containers (host mirror of chromo/polymers.pyx).

`PolymerBase` (polymers.pyx:65-300), `SSWLC` (912-1601) and `Chromatin`
(1821-1886) keep the reference's constructor signatures, attribute names,
dtypes and error behaviour.  The arrays live on the host as numpy (as in the
reference); the energy methods and the MC loop run on the GPU through
`chromo_b200.engine.Engine` -- there is no CPU implementation of them here.
Out of scope: Rouse, SSTWLC, LoopedSSTWLC, DetailedChromatin*.
"""
from __future__ import annotations

import numpy as np

from . import beads
from .util import dss_params

empty_1d = np.empty((0,))
empty_2d = np.empty((0, 0))
mty_2d_int = np.empty((0, 0), dtype=int)


class TransformedObject:
    """Carries the 4x4 homogeneous transformation matrix (polymers.pyx:44-62)."""

    def __init__(self):
        self.transformation_mat = np.identity(4, dtype='d')


class PolymerBase(TransformedObject):
    def __init__(self, name, r=empty_2d, *, log_path="", t3=empty_2d, t2=empty_2d, bead_length=empty_1d,
                 lp=0, states=mty_2d_int, binder_names=empty_1d, chemical_mods=mty_2d_int,
                 chemical_mod_names=empty_1d, max_binders=-1):
        super().__init__()
        self.name = name
        self.r = self._f64(r, "r")
        self.t3 = self._f64(t3, "t3")
        self.t2 = self._f64(t2, "t2")
        self.bead_length = np.ascontiguousarray(bead_length, dtype=np.float64)
        self.states = self._i64(states, "states")
        self.max_binders = int(max_binders)
        self.binder_names = np.asarray(binder_names)
        self.chemical_mods = self._i64(chemical_mods, "chemical_mods")
        self.chemical_mod_names = np.asarray(chemical_mod_names)
        self.num_beads = self.get_num_beads()
        self.fill_missing_arguments()
        self.all_inds = np.arange(0, self.num_beads, 1)
        self.num_binders = self.get_num_binders()
        self.n_binders_p1 = self.num_binders + 1
        self.log_path = log_path
        self.check_binders(self.states, self.binder_names)
        self.lp = lp
        self.required_attrs = np.array(["name", "r", "t3", "t2", "states", "binder_names", "num_binders",
                                        "beads", "num_beads", "lp", "bead_length"])
        self._arrays = np.array(['r', 't3', 't2', 'states', 'chemical_mods', 'bead_length'])
        self._3d_arrays = np.array(['r', 't3', 't2'])
        self._single_values = np.array(["name", "lp", "lt", "bp_wrap"])
        # trial state: equal to the current state between moves (moves.pyx:156-299)
        self.r_trial = self.r.copy()
        self.t3_trial = self.t3.copy()
        self.t2_trial = self.t2.copy()
        self.states_trial = self.states.copy()
        self.last_amp_bead = 0
        self.last_amp_move = 0
        self.direction = np.zeros((3,), dtype='d')
        self.point = np.zeros((3,), dtype='d')
        self.mu_adjust_factor = 1
        self._engine = None

    # ---- dtype / layout rules of the typed memoryviews (polymers.pxd:21-24)
    @staticmethod
    def _f64(a, what):
        a = np.asarray(a)
        if a.size and a.dtype != np.float64:
            raise ValueError(f"Buffer dtype mismatch, expected 'double' for `{what}`")
        return np.ascontiguousarray(a, dtype=np.float64)

    @staticmethod
    def _i64(a, what):
        a = np.asarray(a)
        if a.size and a.dtype != np.int64:
            raise ValueError(f"Buffer dtype mismatch, expected 'long' for `{what}`")
        return np.ascontiguousarray(a, dtype=np.int64)

    def fill_missing_arguments(self):
        """polymers.pyx:302-325."""
        if np.size(self.t3) == 0:
            print("No t3 tangent vectors defined.")
            self.t3 = np.zeros((self.num_beads, 3), dtype='d')
        if np.size(self.t2) == 0:
            print("No t2 tangent vectors defined.")
            self.t2 = np.zeros((self.num_beads, 3), dtype='d')
        if np.size(self.states) == 0:
            print("No states defined.")
            self.states = np.zeros((self.num_beads, 1), dtype=int)
        if np.size(self.binder_names) == 0:
            self.binder_names = np.array(["null_reader"])
        if np.size(self.chemical_mods) == 0:
            print("No chemical modifications defined.")
            self.chemical_mods = np.zeros((self.num_beads, self.states.shape[1]), dtype=int)
        if np.size(self.chemical_mod_names) == 0:
            self.chemical_mod_names = np.array(["null_mod"])

    def check_binders(self, states, binder_names):
        """polymers.pyx:404-425."""
        if states.shape[1] != 0:
            if states.shape[1] != len(binder_names):
                raise ValueError("Each chemical state must be given a name.")
            if states.shape[0] != len(self.r):
                raise ValueError("Initial epigenetic state of wrong length.")

    def check_attrs(self):
        for attr in self.required_attrs:
            if not hasattr(self, str(attr)):
                raise NotImplementedError("Polymer subclass missing required attribute: " + str(attr))
        if self.lp == 0:
            raise ValueError("Specify the persistence length in the subclass of Polymer")

    # ---- snapshot I/O: the reference's CSV schema (polymers.pyx:575-792) ----
    def to_dataframe(self):
        """One row per bead under a two-level header: (r|t3|t2, x|y|z), (states, binder),
        (chemical_mods, mark), then single-level columns bead_length (N-1 bond lengths and a
        trailing 0), the other per-polymer arrays (`max_binders`, broadcast) and the scalars
        (`name`, `lp`, ...: value in row 0, empty below).  Same columns, order and cell text as
        PolymerBase.to_dataframe (polymers.pyx:575-668)."""
        import pandas as pd
        n = self.num_beads
        cols, data = [], []
        held = [str(a) for a in self._arrays if hasattr(self, str(a))]
        for name in held:
            if name in self._3d_arrays:
                arr = np.asarray(getattr(self, name), dtype=np.float64)
                for k, axis in enumerate("xyz"):
                    cols.append((name, axis))
                    data.append(arr[:, k])
        if len(self.chemical_mod_names) > 0:
            st = np.asarray(self.states)
            for j, nm in enumerate(self.binder_names):
                cols.append(("states", str(nm)))
                data.append(st[:, j])
            cm = np.asarray(self.chemical_mods)
            for j, nm in enumerate(self.chemical_mod_names):
                cols.append(("chemical_mods", str(nm)))
                data.append(cm[:, j].astype(int))
        cols.append(("bead_length", ""))
        data.append(np.append(np.asarray(self.bead_length, dtype=np.float64), 0.0))
        for name in held:
            if name in self._3d_arrays or name in ("states", "chemical_mods", "bead_length"):
                continue
            cols.append((name, ""))
            data.append(np.broadcast_to(np.asarray(getattr(self, name)), (n,)))
        for name in self._single_values:
            name = str(name)
            if not hasattr(self, name):
                continue
            col = np.full(n, "", dtype=object)
            val = getattr(self, name)
            col[0] = str(float(val)) if name != "name" else str(val)
            cols.append((name, ""))
            data.append(col)
        df = pd.DataFrame({i: d for i, d in enumerate(data)})
        df.columns = pd.MultiIndex.from_tuples(cols)
        return df

    def to_csv(self, path):
        """polymers.pyx:670-684."""
        return self.to_dataframe().to_csv(path)

    def to_file(self, path):
        """Synonym of `to_csv` (polymers.pyx:686-693)."""
        return self.to_csv(path)

    @classmethod
    def from_dataframe(cls, df, name=None, **kwargs):
        """Inverse of `to_dataframe` (polymers.pyx:713-765): the top-level column names are
        the constructor's keyword arguments."""
        top = list(dict.fromkeys(df.columns.get_level_values(0)))
        kw = {}
        for key in top:
            kw[key] = np.array(df[key].to_numpy(), order="C", copy=True)  # pandas hands out read-only views
        if "states" in top:
            kw["binder_names"] = df["states"].columns.to_numpy()
            kw["states"] = np.array(kw["states"], dtype=np.int64, order="C")
        if "chemical_mods" in top:
            kw["chemical_mod_names"] = df["chemical_mods"].columns.to_numpy()
            kw["chemical_mods"] = np.array(kw["chemical_mods"], dtype=np.int64, order="C")
        for key in ("r", "t3", "t2"):
            if key in kw:
                kw[key] = np.array(kw[key], dtype=np.float64, order="C")
        if "max_binders" in top:
            kw["max_binders"] = int(np.ravel(kw["max_binders"])[0])
        stored = np.ravel(kw.pop("name"))[0] if "name" in top else None
        kw["name"] = name if name is not None else (stored if stored is not None else "unnamed")
        for key in ("lp", "lt", "bp_wrap"):
            if key in top:
                kw[key] = float(np.ravel(kw[key])[0])
        if "bead_length" in top:
            kw["bead_length"] = np.ravel(kw["bead_length"])[:-1].astype(float)
        kw.update(kwargs)
        return cls(**kw)

    @classmethod
    def from_csv(cls, csv_file):
        """polymers.pyx:695-710 (reads a file written by `to_csv`)."""
        return cls.from_file(csv_file)

    @classmethod
    def from_file(cls, path, name=None, **kwargs):
        """polymers.pyx:767-792: the polymer is named after the file unless `name` is given."""
        import pandas as pd
        if name is None:
            name = str(path).split("/")[-1].split(".")[0]
        # round-trip float parsing: a snapshot reloads bit for bit (the reference's default parser can be
        # off by an ulp, which a resumed run would carry along)
        df = pd.read_csv(path, header=[0, 1], index_col=0, float_precision="round_trip")
        return cls.from_dataframe(df, name, **kwargs)

    def update_log_path(self, log_path):
        """polymers.pyx:794-802."""
        self.log_path = log_path

    def get_num_binders(self):
        return self.states.shape[1]

    def get_num_beads(self):
        return self.r.shape[0]

    def is_field_active(self):
        return 1  # polymers.pyx:812-832

    def __str__(self):
        return f"Polymer<{self.name}, nbeads={self.num_beads}, nbinder={self.num_binders}>"


class SSWLC(PolymerBase):
    """Stretchable, shearable wormlike chain (polymers.pyx:912-1601)."""

    _bead_cls = beads.GhostBead

    def __init__(self, name, r, *, bead_length, lp, bead_rad=5, t3=empty_2d, t2=empty_2d,
                 states=mty_2d_int, binder_names=empty_1d, chemical_mods=mty_2d_int,
                 chemical_mod_names=empty_1d, log_path="", max_binders=-1):
        super().__init__(name, r, t3=t3, t2=t2, states=states, binder_names=binder_names,
                         bead_length=bead_length, lp=lp, log_path=log_path, chemical_mods=chemical_mods,
                         chemical_mod_names=chemical_mod_names, max_binders=max_binders)
        self.bead_rad = bead_rad
        self.construct_beads()
        self._find_parameters(self.bead_length)
        self.required_attrs = np.array(["name", "r", "t3", "t2", "states", "binder_names", "num_binders",
                                        "beads", "num_beads", "lp", "bead_rad"])
        self._arrays = np.array(['r', 't3', 't2', 'states', 'bead_length', 'chemical_mods', 'max_binders'])
        self.check_attrs()
        self.mu_adjust_factor = 1

    def construct_beads(self):
        self.beads = beads.BeadMap(self, self._bead_cls)

    def _find_parameters(self, bead_length):
        """Elastic parameters of each bond interpolated from the dssWLC table
        (polymers.pyx:1545-1601).  Bonds of equal length share one lookup."""
        bl = np.asarray(bead_length, dtype=float)
        n = len(bl)
        names = ("delta", "eps_bend", "gamma", "eps_par", "eps_perp", "eta")
        out = {k: np.zeros(n) for k in names}
        cache = {}
        for i in range(n):
            key = bl[i]
            if key not in cache:
                d = bl[i] / self.lp
                cache[key] = (
                    d,
                    np.interp(d, dss_params[:, 0], dss_params[:, 1]) / d,
                    np.interp(d, dss_params[:, 0], dss_params[:, 2]) * d * self.lp,
                    np.interp(d, dss_params[:, 0], dss_params[:, 3]) / (d * self.lp ** 2),
                    np.interp(d, dss_params[:, 0], dss_params[:, 4]) / (d * self.lp ** 2),
                    np.interp(d, dss_params[:, 0], dss_params[:, 5]) / self.lp,
                )
            for k, v in zip(names, cache[key]):
                out[k][i] = v
        for k in names:
            setattr(self, k, out[k])

    # ---- device-backed energies -------------------------------------------
    def _polymer_engine(self):
        """A 1-replica, field-less engine for SSWLC.compute_E."""
        from .engine import Engine
        from .fields import binder_dicts, _zero_prefactors
        if self._engine is None:
            e = Engine(1, self.num_beads, self.num_binders, grid=None, bead_vol=self.beads[0].vol,
                       max_binders=self.max_binders)
            bd = binder_dicts(self)
            e.set_binders(bd, *_zero_prefactors(self.num_binders))
            e.set_bond_params(self.eps_bend, self.eps_par, self.eps_perp, self.gamma, self.eta)
            if getattr(self, "eps_twist", None) is not None:  # SSTWLC
                e.set_twist_params(self.eps_twist, self.natural_twist)
                if getattr(self, "nucleosome_constants", None) is not None:  # DetailedChromatin
                    e.set_detailed_nucleosomes(self.nucleosome_constants)
            self._engine = e
        return self._engine

    def compute_E(self):
        """Total elastic energy (SSWLC.compute_E, polymers.pyx:1348-1381), on the GPU."""
        e = self._polymer_engine()
        e.upload(self.r[None], self.t3[None], self.t2[None], self.states[None], self.chemical_mods[None])
        return float(e.elastic_energy()[0])

    def __str__(self):
        return f"Polymer_Class<SSWLC>, {super().__str__()}"

    # ---- initialisers (polymers.pyx:1603-1818) ------------------------------
    @classmethod
    def straight_line_in_x(cls, name, step_sizes, **kwargs):
        csum = np.cumsum(step_sizes)
        num_beads = len(step_sizes) + 1
        r = np.zeros((num_beads, 3))
        r[1:, 0] = csum.copy()
        t3 = np.zeros((num_beads, 3))
        t3[:, 0] = 1
        t2 = np.zeros((num_beads, 3))
        t2[:, 1] = 1
        return cls(name, r, t3=t3, t2=t2, bead_length=np.asarray(step_sizes, dtype=float), **kwargs)

    @classmethod
    def gaussian_walk_polymer(cls, name, num_beads, step_lengths, **kwargs):
        from .util import poly_paths as paths
        r = paths.gaussian_walk(num_beads - 1, step_lengths, np.random.default_rng(np.random.randint(2 ** 31)))
        t3, t2 = paths.estimate_tangents_from_coordinates(r)
        return cls(name, r, t3=t3, t2=t2, bead_length=np.asarray(step_lengths, dtype=float), **kwargs)

    @classmethod
    def confined_gaussian_walk(cls, name, num_beads, step_lengths, confine_type, confine_length, **kwargs):
        from .util import poly_paths as paths
        step_lengths = np.ascontiguousarray(step_lengths, dtype=float)
        r = paths.confined_gaussian_walk(num_beads, step_lengths, confine_type, confine_length,
                                         np.random.default_rng(np.random.randint(2 ** 31)))
        t3, t2 = paths.estimate_tangents_from_coordinates(r)
        return cls(name, r, t3=t3, t2=t2, bead_length=step_lengths, **kwargs)


LENGTH_BP = 0.332                      # polymers.pyx:38
NATURAL_TWIST_BARE = 2 * np.pi / 10.5  # polymers.pyx:40


class SSTWLC(SSWLC):
    """Stretchable, shearable wormlike chain with twist (polymers.pyx:1889-2319): every bond energy
    carries 0.5 * eps_twist * wrap(omega - natural twist)^2 (E_pair_with_twist 2050-2102), in the MC
    kernel (`-DCB_TWIST` builds) and in `compute_E`."""

    def __init__(self, name, r, *, bead_length, lp, lt, bead_rad=5, t3=empty_2d, t2=empty_2d,
                 states=mty_2d_int, binder_names=empty_1d, chemical_mods=mty_2d_int,
                 chemical_mod_names=empty_1d, log_path="", max_binders=-1):
        self.lt = float(lt)
        super().__init__(name, r, bead_length=bead_length, lp=lp, bead_rad=bead_rad, t3=t3, t2=t2, states=states,
                         binder_names=binder_names, chemical_mods=chemical_mods,
                         chemical_mod_names=chemical_mod_names, log_path=log_path, max_binders=max_binders)
        self.required_attrs = np.array(["name", "r", "t3", "t2", "states", "binder_names", "num_binders",
                                        "beads", "num_beads", "lp", "lt", "bead_rad"])
        self._arrays = np.array(['r', 't3', 't2', 'states', 'bead_length', 'chemical_mods'])
        self.check_attrs()

    def _find_parameters(self, bead_length):
        """polymers.pyx:1957-2001: the SSWLC parameters plus eps_twist = lt / (delta * lp); the natural
        twist of a bond is bead_length * NATURAL_TWIST_BARE / LENGTH_BP (2088-2090)."""
        super()._find_parameters(bead_length)
        self.eps_twist = self.lt / (self.delta * self.lp)
        self.natural_twist = np.asarray(bead_length, dtype=float) * NATURAL_TWIST_BARE / LENGTH_BP

    def compute_E_no_twist(self):
        """polymers.pyx:2287-2319: the elastic energy without the twist term."""
        e = self._polymer_engine()
        e.set_twist_params(None)
        try:
            return super().compute_E()
        finally:
            e.set_twist_params(self.eps_twist, self.natural_twist)

    def __str__(self):
        return f"Polymer_Class<SSTWLC>, {PolymerBase.__str__(self)}"


class DetailedChromatin(SSTWLC):
    """Chromatin fiber with detailed nucleosomes (polymers.pyx:2455-2607): an SSTWLC whose elastic dE runs every
    linker from the EXIT point / frame of one nucleosome to the ENTRY point / frame of the next
    (DetailedNucleosome, beads.py:448-574).  `bead_rad` is the nucleosome's radius (nucleo_geom R); `compute_E`
    stays the SSTWLC one, as in the reference (only `continuous_dE_poly` is overridden there)."""

    _with_diameter = True

    def __init__(self, name, r, *, bp_wrap, bead_length, lp, lt, t3=empty_2d, t2=empty_2d, states=mty_2d_int,
                 binder_names=empty_1d, chemical_mods=mty_2d_int, chemical_mod_names=empty_1d, log_path="",
                 max_binders=-1):
        from .util import nucleo_geom
        self.bp_wrap = float(bp_wrap)
        self.nucleosome_constants = nucleo_geom.nucleosome_constants(self.bp_wrap, self._with_diameter)
        super().__init__(name, r, bead_length=bead_length, lp=lp, lt=lt, bead_rad=nucleo_geom.consts_dict["R"], t3=t3,
                         t2=t2, states=states, binder_names=binder_names, chemical_mods=chemical_mods,
                         chemical_mod_names=chemical_mod_names, log_path=log_path, max_binders=max_binders)

    def __str__(self):
        return f"Polymer_Class<DetailedChromatin>, {PolymerBase.__str__(self)}"


class DetailedChromatin2(DetailedChromatin):
    """`DetailedChromatin` without the nucleosome diameter (polymers.pyx:2627-2735): the linker runs between the
    bead centres, only the exit frame of the nucleosome enters the bond energy (the kinked wormlike chain)."""
    _with_diameter = False

    def __str__(self):
        return f"Polymer_Class<DetailedChromatin2>, {PolymerBase.__str__(self)}"


class Chromatin(SSWLC):
    """SSWLC model of chromatin, lp = 53 nm (polymers.pyx:1821-1886)."""

    _bead_cls = beads.Nucleosome

    def __init__(self, name, r, *, bead_length, bead_rad=5, t3=empty_2d, t2=empty_2d, states=mty_2d_int,
                 binder_names=empty_1d, chemical_mods=mty_2d_int, chemical_mod_names=empty_1d,
                 log_path="", max_binders=-1, **kwargs):
        for nm, a in (("r", r), ("t3", t3), ("t2", t2), ("states", states), ("chemical_mods", chemical_mods)):
            if not isinstance(a, np.ndarray):  # buffer-typed arguments, polymers.pyx:1842-1851
                raise TypeError(f"Argument '{nm}' has incorrect type (expected numpy.ndarray, got {type(a).__name__})")
        super().__init__(name, r, bead_length=bead_length, bead_rad=bead_rad, lp=53, t3=t3, t2=t2,
                         states=states, binder_names=binder_names, log_path=log_path,
                         chemical_mods=chemical_mods, chemical_mod_names=chemical_mod_names,
                         max_binders=max_binders)
