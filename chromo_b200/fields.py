"""Fields (host mirror of chromo/fields.pyx): `FieldBase` (46-202), `NullField`
(280-321) and `UniformDensityField` (324-2315).

Constructor signatures, attribute names and error behaviour follow the
reference.  The voxel densities live in HBM; `field.density` is a host copy
refreshed after every device call that changes it.  All energy evaluation
happens in the CUDA kernels (chromo_b200/csrc) -- nothing here computes a dE.
`fast_field=1` (sub-bin quantised binning in the dE path, fields.pyx:577-671, 1235-1368) is a
kernel option (`chromo_ctx_set_fast_field`).  Out of scope: Reconstructor, neighbour-bin helpers.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

_field_descriptors = ['x_width', 'nx', 'y_width', 'ny', 'z_width', 'nz', 'confine_type', 'confine_length',
                      'chi', 'assume_fully_accessible', 'vf_limit', 'fast_field']
_int_field_descriptors = ['nx', 'ny', 'nz', 'n_points']
_str_field_descriptors = ['confine_type']
_float_field_descriptors = ['x_width', 'y_width', 'z_width', 'confine_length', 'vf_limit', 'chi']
_bool_field_descriptors = ['assume_fully_accessible', 'fast_field']


def binder_dicts(poly):
    """Per-binder parameters as the hot path reads them: from the LIVE binder
    singletons resolved by name at bead level (beads.py:74-76)."""
    out = []
    for b in poly.beads[0].binders:
        out.append(dict(name=b.name, sites_per_bead=int(b.sites_per_bead),
                        bind_energy_mod=float(b.bind_energy_mod),
                        bind_energy_no_mod=float(b.bind_energy_no_mod),
                        chemical_potential=float(b.chemical_potential)))
    return out


def _zero_prefactors(nb):
    return np.zeros(nb), np.zeros(nb), np.zeros((nb, nb))


def accessible_volumes(grid, n_side=20, assume_fully_accessible=1):
    """Per-voxel accessible volume of a grid description (get_accessible_volumes fields.pyx:714-770).
    With assume_fully_accessible == 0 and a spherical confinement, voxels cut by the sphere get
    vol_bin * (fraction of an n_side^3 sub-grid, anchored at the voxel corner, that lies inside)."""
    nx, ny, nz = int(grid["nx"]), int(grid["ny"]), int(grid["nz"])
    n_bins = nx * ny * nz
    dx, dy, dz = grid["x_width"] / nx, grid["y_width"] / ny, grid["z_width"] / nz
    vol_bin = grid["x_width"] * grid["y_width"] * grid["z_width"] / n_bins
    vols = np.full(n_bins, vol_bin)
    if assume_fully_accessible == 1 or grid.get("confine_type", "") != "Spherical":
        return vols
    R = float(grid["confine_length"])
    ii = np.arange(n_bins)
    ix, iy, iz = ii % nx, (ii // nx) % ny, ii // (nx * ny)
    # voxel "centres", get_voxel_coords fields.pyx:772-803.  `(n - 1) / 2` there divides two C longs with
    # Python semantics, i.e. FLOORS: on an even grid the reference's centres sit half a voxel below the
    # geometric ones, and the accessible volumes follow from those (reproduced, SURVEY quirk list)
    centre = np.stack([(ix - (nx - 1) // 2) * dx, (iy - (ny - 1) // 2) * dy, (iz - (nz - 1) // 2) * dz], axis=1)
    # voxels cut by the sphere, get_split_voxels fields.pyx:805-840
    buffer_dist = np.sqrt(2) / 4 * max(dx, dy, dz)
    dist = np.sqrt(centre[:, 0] ** 2 + centre[:, 1] ** 2 + centre[:, 2] ** 2)
    split = ~((dist < R - buffer_dist) | (dist > R + buffer_dist))
    # n_side^3 sub-grid anchored at the voxel corner, define_voxel_subgrid 842-880
    k = np.arange(n_side, dtype=float)
    sub = np.stack(np.meshgrid(k * (dx / n_side), k * (dy / n_side), k * (dz / n_side), indexing="ij"),
                   -1).reshape(-1, 3)
    corner = centre - np.array([dx / 2, dy / 2, dz / 2])
    for b in np.nonzero(split)[0]:  # get_frac_accessible fields.pyx:882-951
        inside = np.linalg.norm(corner[b] + sub, axis=1) < R
        vols[b] = vol_bin * (inside.sum() / float(len(sub)))
    return vols


class FieldBase:
    """A discretization of space for computing energies (fields.pyx:46-202)."""

    @property
    def name(self):
        return self.__class__.__name__

    def __init__(self, polymers, binders):
        self.polymers = polymers
        self.n_polymers = len(polymers)
        self.binders = binders
        self.confine_type = ""
        self.confine_length = 0.0
        self._engine = None

    def __str__(self):
        return "Field<>"

    def __contains__(self, poly):
        return self.polymers.__contains__(poly)

    # ---- device plumbing shared by NullField / UniformDensityField ---------
    def _grid(self):
        """Engine grid description; NullField has no voxels but may confine."""
        return dict(nx=0, ny=0, nz=0, x_width=0.0, y_width=0.0, z_width=0.0,
                    confine_type=self.confine_type, confine_length=self.confine_length, vf_limit=0.5)

    def _prefactors(self, nb):
        return _zero_prefactors(nb)

    def _engine_for(self, poly):
        from .engine import Engine
        if self._engine is None or self._engine_poly is not poly:
            if self.confine_type not in ("", "Spherical", "Cubical"):
                raise ValueError("Confinement type " + self.confine_type + " not found.")
            e = Engine(1, poly.num_beads, poly.num_binders, grid=self._grid(), bead_vol=poly.beads[0].vol,
                       max_binders=poly.max_binders)
            e.set_binders(binder_dicts(poly), *self._prefactors(poly.num_binders))
            e.set_bond_params(poly.eps_bend, poly.eps_par, poly.eps_perp, poly.gamma, poly.eta)
            if getattr(poly, "eps_twist", None) is not None:  # SSTWLC
                e.set_twist_params(poly.eps_twist, poly.natural_twist)
                if getattr(poly, "nucleosome_constants", None) is not None:  # DetailedChromatin
                    e.set_detailed_nucleosomes(poly.nucleosome_constants)
            self._engine, self._engine_poly = e, poly
        return self._engine

    def _push(self, poly, with_density=True):
        e = self._engine_for(poly)
        e.upload(poly.r[None], poly.t3[None], poly.t2[None], poly.states[None], poly.chemical_mods[None])
        e.set_replica_params(chi=getattr(self, "chi", 1.0),
                             mu=[b["chemical_potential"] for b in binder_dicts(poly)])
        if with_density and e.n_bins:
            e.upload_density(np.ascontiguousarray(self.density)[None])
        return e

    def _pull(self, poly, e):
        r, t3, t2, st = e.download()
        poly.r[...], poly.t3[...], poly.t2[...], poly.states[...] = r[0], t3[0], t2[0], st[0]
        poly.r_trial[...], poly.t3_trial[...], poly.t2_trial[...] = r[0], t3[0], t2[0]
        poly.states_trial[...] = st[0]
        if e.n_bins:
            self.density[...] = e.density()[0]


class NullField(FieldBase):
    """No density field; only the confinement acts (fields.pyx:280-321)."""

    def __init__(self, polymers=None, confine_type="", confine_length=0.0):
        super().__init__([] if polymers is None else polymers, binders=pd.DataFrame())
        self.confine_type = confine_type
        self.confine_length = confine_length

    def to_file(self, path):
        with open(path, 'w'):
            pass

    @classmethod
    def from_file(cls, path):
        return cls()


class UniformDensityField(FieldBase):
    """Rectilinear voxel grid with trilinear density interpolation (fields.pyx:324-2315)."""

    def __init__(self, polymers, binders, x_width, nx, y_width, ny, z_width, nz, confine_type="",
                 confine_length=0.0, chi=1.0, assume_fully_accessible=1, vf_limit=0.5, fast_field=0,
                 n_points=1000):
        super().__init__(polymers=polymers, binders=binders)
        self._field_descriptors = _field_descriptors
        for poly in polymers:
            if poly.num_binders != len(binders):
                raise NotImplementedError("For now, all polymers must use all of the same binders.")
        if len(polymers) != 1:
            raise NotImplementedError("chromo_b200 evaluates one polymer per field (fields.pyx:53-59); "
                                      "use chromo_b200.ensemble.ReplicaEnsemble for many replicas.")
        self.x_width, self.y_width, self.z_width = float(x_width), float(y_width), float(z_width)
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.init_grid()
        self.num_binders = len(binders)
        self.doubly_bound = np.zeros((self.num_binders,), dtype=int)
        self.doubly_bound_trial = np.zeros((self.num_binders,), dtype=int)
        self.init_field_energy_prefactors()
        self.density = np.zeros((self.n_bins, self.num_binders + 1), dtype=float)
        self.density_trial = self.density.copy()
        self.confine_type = confine_type
        self.confine_length = float(confine_length)
        self.assume_fully_accessible = assume_fully_accessible
        self.chi = float(chi)
        self.vf_limit = float(np.float32(vf_limit))  # C float in the reference (fields.pxd:61)
        self.access_vols = self.get_accessible_volumes(n_side=20, assume_fully_accessible=assume_fully_accessible)
        self.binder_dict = self.binders.to_dict(orient='records')
        self.fast_field = fast_field
        self.n_points = n_points
        self.affected_bins_last_move = np.zeros((self.n_bins,), dtype=int)
        self.update_all_densities_for_all_polymers()
        self.dict_ = self.get_dict()

    # ---- grid (fields.pyx:536-575) ------------------------------------------
    def init_grid(self):
        self.dx = self.x_width / self.nx
        self.dy = self.y_width / self.ny
        self.dz = self.z_width / self.nz
        self.dxyz = np.array([self.dx, self.dy, self.dz])
        self.n_bins = self.nx * self.ny * self.nz
        self.vol_bin = self.x_width * self.y_width * self.z_width / self.n_bins
        self.width_xyz = np.array([self.x_width, self.y_width, self.z_width])
        self.half_width_xyz = 0.5 * self.width_xyz
        self.half_step_xyz = np.array([0.5 * self.dx, 0.5 * self.dy, 0.5 * self.dz])
        self.n_xyz_m1 = np.array([self.nx - 1, self.ny - 1, self.nz - 1])

    def init_field_energy_prefactors(self):
        """fields.pyx:687-712; like the reference, writes into the binders table."""
        names = [self.binders.loc[i, "name"] for i in range(self.num_binders)]
        for i in range(self.num_binders):
            row = self.binders.iloc[i]
            self.binders.at[i, 'field_energy_prefactor'] = (
                0.5 * row.interaction_energy * row.interaction_volume * self.vol_bin)
            self.binders.at[i, 'interaction_energy_intranucleosome'] = (
                row.interaction_energy * (1 - row.interaction_volume / self.vol_bin))
            for nxt in names:
                if nxt in row.cross_talk_interaction_energy.keys():
                    self.binders.at[i, 'cross_talk_field_energy_prefactor'][nxt] = (
                        row.cross_talk_interaction_energy[nxt] * row.interaction_volume * self.vol_bin)
                else:
                    self.binders.at[i, 'cross_talk_field_energy_prefactor'][nxt] = 0

    def get_accessible_volumes(self, n_side, assume_fully_accessible):
        """Per-voxel accessible volume (fields.pyx:714-770)."""
        return accessible_volumes(self._grid_geometry(), n_side, assume_fully_accessible)

    def _grid_geometry(self):
        return dict(nx=self.nx, ny=self.ny, nz=self.nz, x_width=self.x_width, y_width=self.y_width,
                    z_width=self.z_width, confine_type=self.confine_type, confine_length=self.confine_length)

    def get_dict(self):
        return {"x_width": self.x_width, "y_width": self.y_width, "z_width": self.z_width, "nx": self.nx,
                "ny": self.ny, "nz": self.nz, "num_binders": self.num_binders, "n_bins": self.n_bins,
                "density": self.density, "confine_type": self.confine_type,
                "confine_length": self.confine_length, "chi": self.chi,
                "assume_fully_accessible": self.assume_fully_accessible, "vf_limit": self.vf_limit,
                "fast_field": self.fast_field, "n_points": self.n_points}

    def __str__(self):
        return f"UniformDensityField<nx={self.nx},ny={self.ny},nz={self.nz}>"

    # ---- device plumbing -----------------------------------------------------
    def _grid(self):
        return dict(nx=self.nx, ny=self.ny, nz=self.nz, x_width=self.x_width, y_width=self.y_width,
                    z_width=self.z_width, confine_type=self.confine_type,
                    confine_length=self.confine_length, vf_limit=self.vf_limit)

    def _prefactors(self, nb):
        bd = self.binder_dict
        pref = np.array([b["field_energy_prefactor"] for b in bd], dtype=float)
        e_intra = np.array([b["interaction_energy_intranucleosome"] for b in bd], dtype=float)
        xpref = np.array([[b["cross_talk_field_energy_prefactor"].get(c["name"], 0) for c in bd] for b in bd],
                         dtype=float).reshape(nb, nb)
        return pref, e_intra, xpref

    def _engine_for(self, poly):
        fresh = self._engine is None or self._engine_poly is not poly
        e = super()._engine_for(poly)
        if fresh and self.assume_fully_accessible != 1:
            e.set_access_volumes(self.access_vols)
        if fresh and self.fast_field == 1:  # sub-bin quantised binning in the dE path (fields.pyx:577-671, 1235-1368)
            e.set_fast_field(self.n_points)
        return e

    # ---- full recompute / total energy (A8) ----------------------------------
    def update_all_densities_for_all_polymers(self):
        """fields.pyx:2041-2106 (scatter + |rho| < 1e-18 clamp), on the GPU."""
        poly = self.polymers[0]
        e = self._push(poly, with_density=False)
        e.field_recompute(clamp=True)
        self.density[...] = e.density()[0]
        self.density_trial[...] = 0

    def update_all_densities(self, poly, inds=None, n_inds=None):
        """fields.pyx:1977-2039, on the GPU."""
        e = self._push(poly, with_density=False)
        e.field_recompute(clamp=False)
        self.density[...] = e.density()[0]
        self.density_trial[...] = 0

    def compute_E(self, poly):
        """Total field energy (fields.pyx:1939-1966, 2208-2315), on the GPU."""
        e = self._push(poly, with_density=False)
        E, _, dbl, _ = e.field_energy()
        self.density[...] = e.density()[0]
        self.density_trial[...] = 0
        self.doubly_bound[...] = dbl[0]
        return float(E[0])

    def nonspecific_interact_E(self, poly):
        e = self._push(poly, with_density=True)
        e.field_recompute(clamp=False)
        return float(e.field_energy()[3][0])

    # ---- CSV round trip (fields.pyx:953-1036) ---------------------------------
    def to_file(self, path):
        rows = {name: self.dict_[name] for name in self._field_descriptors}
        for polymer in self.polymers:
            rows[polymer.name] = 'polymer'
        for _, binder in self.binders.iterrows():
            rows[binder['name']] = 'binder'
        return pd.Series(rows).to_csv(path, header=False)

    @classmethod
    def from_file(cls, path, polymers, binders):
        field_series = pd.read_csv(path, header=None, index_col=0)[1]
        kwargs = field_series[_field_descriptors].to_dict()
        for key in kwargs.keys():
            if key in _int_field_descriptors:
                kwargs[key] = int(kwargs[key])
            elif key in _float_field_descriptors:
                kwargs[key] = float(kwargs[key])
            elif key in _str_field_descriptors and pd.isna(kwargs[key]):
                kwargs[key] = ""
            elif key in _bool_field_descriptors:
                kwargs[key] = int(str(kwargs[key]) in ("1", "True", "1.0"))
        polymer_names = field_series[field_series == 'polymer'].index.values
        binder_names = field_series[field_series == 'binder'].index.values
        err_prefix = f"Tried to instantiate class:{cls} from file:{path} with "
        if len(polymers) != len(polymer_names):
            raise ValueError(err_prefix + f"{len(polymers)} polymers, but  there are {len(polymer_names)} listed.")
        for polymer in polymers:
            if polymer.name not in polymer_names:
                raise ValueError(err_prefix + f"polymer:{polymer.name}, but  this polymer was not present in file.")
        if len(binders) != len(binder_names):
            raise ValueError(err_prefix + f"{len(binders)} binders, but  there are {len(binder_names)} listed.")
        for _, binder in binders.iterrows():
            if binder['name'] not in binder_names:
                raise ValueError(err_prefix + f"binder:{binder}, but  this binder was not present in file.")
        return cls(polymers=polymers, binders=binders, **kwargs)

    def __eq__(self, other):
        return all(self.dict_[k] == other.dict_[k] for k in self._field_descriptors)

    __hash__ = object.__hash__
