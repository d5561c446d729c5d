"""ReplicaEnsemble: many independent replicas / parameter points of one problem
shape, resident in the HBM of one B200 and advanced by ONE kernel launch per
`mc_sim` call (one warp per replica).

This is the batched form of `chromo_b200.mc.mc_sim`: the reference runs one
simulation per process; chi / chemical-potential sweeps and
replica-exchange ladders are many such simulations, which is exactly what a GPU
wants.  Replicas shard across GPUs by `chromo_b200.parallel`.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import MOVE_DTYPE, MOVE_NAMES, NUM_MOVES, RNG_PHILOX, RNG_REPLAY
from .engine import Engine


def default_moves(n_replicas: int, num_beads: int, min_spacing: float, controller: int = 1,
                  per_cycle: Sequence[int] = (30, 1, 60, 60, 10), move_on: Sequence[int] = (1, 1, 1, 1, 1)):
    """[R,5] controller state as `all_moves(..., SimpleControl)` builds it from
    `get_amplitude_bounds` (mc/__init__.py:295-332, mc_controller.py:216-266)."""
    N = num_beads
    bead = [(min(30, N), min(150, N)), (min(50, N / 4), min(150, int(N / 2))), (min(10, N), min(150, N)),
            (1, N), (1, 1)]
    move = [(0.1 * np.pi, 0.25 * np.pi), (0.2 * np.pi, 0.25 * np.pi), (0.2 * min_spacing, 0.3 * min_spacing),
            (0.05 * np.pi, 0.2 * np.pi), (0, 0)]
    a = np.zeros((n_replicas, NUM_MOVES), dtype=MOVE_DTYPE)
    for i in range(NUM_MOVES):
        a["amp_move"][:, i] = move[i][0]
        a["move_amp_lo"][:, i], a["move_amp_hi"][:, i] = move[i]
        a["bead_amp_lo"][:, i], a["bead_amp_hi"][:, i] = bead[i]
        a["amp_bead"][:, i] = int(bead[i][0])
        a["alpha"][:, i] = 2 / (20.0 + 1)
        a["num_per_cycle"][:, i] = per_cycle[i]
        a["move_on"][:, i] = move_on[i]
        a["controller"][:, i] = controller
    return a


class ReplicaEnsemble:
    """R replicas: positions / orientations / states / marks as [R,N,.] host
    arrays, one grid description, per-replica chi and chemical potentials."""

    def __init__(self, r, t3, t2, states, chemical_mods, *, binders: Sequence[dict], bond_params: dict,
                 grid: Optional[dict], bead_vol: float, chi=1.0, mu=None, max_binders: int = -1,
                 moves: Optional[np.ndarray] = None, min_spacing: Optional[float] = None,
                 access_vol=None, device: int = 0, field_prefactors=None, assume_fully_accessible: int = 1,
                 replica_offset: int = 0, fast_field_points: int = 0, pin_host: Optional[bool] = None):
        self.r = np.ascontiguousarray(r, dtype=np.float64)
        self.R, self.N = self.r.shape[0], self.r.shape[1]
        self.t3 = np.ascontiguousarray(t3, dtype=np.float64)
        self.t2 = np.ascontiguousarray(t2, dtype=np.float64)
        self.states = np.ascontiguousarray(states, dtype=np.int64).reshape(self.R, self.N, -1)
        self.nb = self.states.shape[2]
        self.chemical_mods = np.ascontiguousarray(chemical_mods, dtype=np.int64).reshape(self.R, self.N, self.nb)
        self.binders = [dict(b) for b in binders]
        self._pinned = []
        self.grid = grid
        self.bead_vol, self.max_binders, self.bond_params = bead_vol, max_binders, dict(bond_params)
        self.min_spacing, self.device = min_spacing, device
        self.engine = Engine(self.R, self.N, self.nb, grid=grid, bead_vol=bead_vol, max_binders=max_binders,
                             device=device)
        if field_prefactors is None:
            field_prefactors = self.prefactors_from_binders(self.binders, grid)
        self.engine.set_binders(self.binders, *field_prefactors)
        self.engine.set_bond_params(bond_params["eps_bend"], bond_params["eps_par"], bond_params["eps_perp"],
                                    bond_params["gamma"], bond_params["eta"])
        if bond_params.get("eps_twist") is not None:  # SSTWLC replicas (polymers.pyx:1889-2319)
            self.engine.set_twist_params(bond_params["eps_twist"], bond_params["natural_twist"])
            if bond_params.get("nucleosome_constants") is not None:  # DetailedChromatin replicas (polymers.pyx:2455-2607)
                self.engine.set_detailed_nucleosomes(bond_params["nucleosome_constants"])
        self.assume_fully_accessible = assume_fully_accessible
        if access_vol is None and assume_fully_accessible != 1 and grid is not None and grid.get("nx", 0):
            from .fields import accessible_volumes
            access_vol = accessible_volumes(grid, 20, assume_fully_accessible)  # fields.pyx:714-770
        if access_vol is not None:
            self.engine.set_access_volumes(access_vol)
        self.fast_field_points = int(fast_field_points)
        if self.fast_field_points:
            self.engine.set_fast_field(self.fast_field_points)
        self.replica_offset = int(replica_offset)
        if replica_offset:
            self.engine.set_replica_offset(replica_offset)
        self._host_stale = False  # the device state is ahead of the host arrays (sync_host=False calls)
        self.chi = np.ascontiguousarray(np.broadcast_to(np.asarray(chi, dtype=float), (self.R,)))
        if mu is None:
            mu = [b["chemical_potential"] for b in self.binders]
        self.mu = np.ascontiguousarray(np.broadcast_to(np.asarray(mu, dtype=float), (self.R, self.nb)))
        self.engine.set_replica_params(self.chi, self.mu)
        if moves is None:
            moves = default_moves(self.R, self.N, 16.5 if min_spacing is None else min_spacing)
        self.moves = np.ascontiguousarray(moves, dtype=MOVE_DTYPE).reshape(self.R, NUM_MOVES)
        self.engine.set_moves(self.moves)
        # host-in / host-out calls (mc_sim(sync_host=True), push, pull) copy these arrays every time: page-lock them
        # once (pageable numpy memory moves at a third of the link speed).  Default: ensembles of >= 8 MiB.
        if pin_host or (pin_host is None and self.r.nbytes >= (8 << 20)):
            from ._lib import ChromoError
            from .engine import host_register
            try:
                for a in (self.r, self.t3, self.t2, self.states, self.chemical_mods):
                    if host_register(a):
                        self._pinned.append(a)
            except ChromoError:  # e.g. the locked-memory limit: the copies still work, through the driver's staging
                if pin_host:
                    raise
        self.push()
        if grid is not None and grid.get("nx", 0):
            self.engine.field_recompute(clamp=True)  # UniformDensityField.__init__ fields.pyx:532

    @staticmethod
    def prefactors_from_binders(binders, grid):
        """init_field_energy_prefactors fields.pyx:687-712."""
        nb = len(binders)
        pref, e_intra, xpref = np.zeros(nb), np.zeros(nb), np.zeros((nb, nb))
        if grid is None or not grid.get("nx", 0):
            return pref, e_intra, xpref
        vol_bin = grid["x_width"] * grid["y_width"] * grid["z_width"] / (grid["nx"] * grid["ny"] * grid["nz"])
        for i, b in enumerate(binders):
            v_int = b.get("interaction_volume", None)
            if v_int is None:
                v_int = (4.0 / 3.0) * np.pi * b["interaction_radius"] ** 3
            pref[i] = 0.5 * b["interaction_energy"] * v_int * vol_bin
            e_intra[i] = b["interaction_energy"] * (1 - v_int / vol_bin)
            for j, nxt in enumerate(binders):
                if nxt["name"] in b.get("cross_talk", {}):
                    xpref[i, j] = b["cross_talk"][nxt["name"]] * v_int * vol_bin
        return pref, e_intra, xpref

    # ---- host <-> device -------------------------------------------------
    def push(self):
        """Upload the host arrays (user-visible state) to the device."""
        self.engine.upload(self.r, self.t3, self.t2, self.states, self.chemical_mods)
        self._host_stale = False
        self._mods_resident = True  # mc_sim never modifies chemical_mods: later host-array calls leave them on the device

    def pull(self):
        """Refresh the host arrays from the device."""
        self.engine.download_into(self.r, self.t3, self.t2, self.states)
        self._host_stale = False

    def marks_changed(self):
        """Call after editing `chemical_mods` on the host: the next host-array `mc_sim` uploads them again."""
        self._mods_resident = False

    def set_params(self, chi=None, mu=None):
        if chi is not None:
            self.chi = np.ascontiguousarray(np.broadcast_to(np.asarray(chi, dtype=float), (self.R,)))
        if mu is not None:
            self.mu = np.ascontiguousarray(np.broadcast_to(np.asarray(mu, dtype=float), (self.R, self.nb)))
        self.engine.set_replica_params(self.chi, self.mu)

    # ---- the hot path -------------------------------------------------------
    def mc_sim(self, num_mc_steps: int, mu_adjust_factor: float = 1.0, random_seed: int = 0,
               rng: str = "philox", sync_host: bool = True, numpy_seeds=None, n_chunks: int = 0):
        """`mc_sim` (mc_sim.pyx:26-103) for every replica.  With sync_host the
        call is host-in / host-out like the reference's: the host arrays go to the
        device, the kernel runs and they come back, pipelined over `n_chunks` replica
        chunks (0 = automatic; -1 = the unpipelined upload / run / download sequence).
        Without it the state stays in HBM and the call returns as soon as the kernel is
        queued."""
        mode = {"philox": RNG_PHILOX, "replay": RNG_REPLAY}[rng]
        ns = ((random_seed if numpy_seeds is None else numpy_seeds) if mode == RNG_REPLAY else None)
        if sync_host and self._host_stale:
            # device-resident calls ran since the host arrays were last refreshed: the host-array call
            # below would upload the OLD configuration over the advanced one (and its density)
            self.engine.sync()
            self.pull()
            self.moves = self.engine.get_moves()
        if sync_host and n_chunks >= 0:
            # the marks go up again only after `marks_changed()` (the reference never writes them either)
            self.engine.mc_sim_host(num_mc_steps, self.r, self.t3, self.t2, self.states,
                                    None if self._mods_resident else self.chemical_mods,
                                    self.moves, mu_adjust_factor, random_seed, mode, numpy_seeds=ns,
                                    n_chunks=n_chunks)
            self._mods_resident = True
        elif sync_host:
            self.push()
            self.engine.mc_sim(num_mc_steps, self.moves, mu_adjust_factor, random_seed, mode, numpy_seeds=ns)
            self.pull()
        else:
            self.engine.mc_sim(num_mc_steps, None, mu_adjust_factor, random_seed, mode,
                               numpy_seeds=numpy_seeds if mode == RNG_REPLAY else None)
            self._host_stale = True

    def sync(self):
        self.engine.sync()
        self.moves = self.engine.get_moves()

    def density(self):
        return self.engine.density()

    def field_energy(self):
        return self.engine.field_energy()[0]

    def elastic_energy(self):
        return self.engine.elastic_energy()

    def acceptance(self):
        m = self.moves
        with np.errstate(invalid="ignore", divide="ignore"):
            return {n: m["num_success"][:, i].sum() / max(1, m["num_attempt"][:, i].sum())
                    for i, n in enumerate(MOVE_NAMES)}

    # ---- batched snapshots ---------------------------------------
    def save_snapshot(self, path, pull: bool = True):
        """Write the whole replica batch as ONE uncompressed .npz (a straight dump of the
        [R,N,.] arrays: r, t3, t2, states, chemical_mods, the move/controller state and the
        per-replica chi / mu).  A reference-style run writes one CSV per polymer and
        snapshot (mc/__init__.py:139-142); at 1,024 replicas x 10,000 beads that is ~2 GB of
        text per snapshot against 0.8 GB here, and `replica_to_csv` still exports any replica
        in the reference's schema."""
        if pull:
            self.engine.sync()
            self.pull()
            self.moves = self.engine.get_moves()
        np.savez(path, r=self.r, t3=self.t3, t2=self.t2, states=self.states, chemical_mods=self.chemical_mods,
                 moves=self.moves.view(np.uint8).reshape(self.R, -1), chi=self.chi, mu=self.mu,
                 rng_counters=self.engine.rng_counters())

    def load_snapshot(self, path):
        """Restore the state written by `save_snapshot` (same R, N, nb) and upload it; the
        densities are recomputed from the loaded configuration."""
        z = np.load(path, allow_pickle=False)
        if z["r"].shape != self.r.shape or z["states"].shape != self.states.shape:
            raise ValueError(f"snapshot holds {z['r'].shape[0]} x {z['r'].shape[1]} beads x "
                             f"{z['states'].shape[2]} binders, the ensemble {self.R} x {self.N} x {self.nb}")
        self.r[...], self.t3[...], self.t2[...] = z["r"], z["t3"], z["t2"]
        self.states[...], self.chemical_mods[...] = z["states"], z["chemical_mods"]
        self.moves = np.ascontiguousarray(z["moves"]).view(MOVE_DTYPE).reshape(self.R, NUM_MOVES).copy()
        self.engine.set_moves(self.moves)
        self.set_params(chi=z["chi"], mu=z["mu"])
        if "rng_counters" in z.files:  # continue the production streams instead of replaying them
            self.engine.set_rng_counters(z["rng_counters"])
        self.push()
        if self.grid is not None and self.grid.get("nx", 0):
            self.engine.field_recompute(clamp=True)

    def replica_polymer(self, i: int, name: Optional[str] = None, *, bead_length, chemical_mod_names=None,
                        bead_rad: float = 5.0):
        """Replica `i` as a `chromo_b200.polymers.Chromatin` (host arrays as of the last pull)."""
        from .polymers import Chromatin
        names = np.array([b["name"] for b in self.binders])
        if chemical_mod_names is None:
            chemical_mod_names = np.array([f"mod{j}" for j in range(self.nb)])
        return Chromatin(name or f"Chr-{i + 1}", self.r[i].copy(), bead_length=np.asarray(bead_length, dtype=float),
                         bead_rad=bead_rad, t3=self.t3[i].copy(), t2=self.t2[i].copy(),
                         states=self.states[i].copy(), binder_names=names,
                         chemical_mods=self.chemical_mods[i].copy(),
                         chemical_mod_names=np.asarray(chemical_mod_names))

    def replica_to_csv(self, i: int, path, **kwargs):
        """Replica `i` in the reference's CSV schema (polymers.pyx:575-684)."""
        return self.replica_polymer(i, **kwargs).to_csv(str(path))

    # ---- coarse-grain / refine (chromo/util/rediscretize.py) -------------
    def _uniform_bonds(self, n_bonds):
        """Bond parameters of a re-discretised chain: the reference gives every bond of the new chain
        the first bond's length (rediscretize.py:447-449, 1059-1060), so uniform parameters carry over."""
        out = {}
        for k, v in self.bond_params.items():
            if v is None:
                continue
            if k == "nucleosome_constants":  # per model, not per bond
                out[k] = v
                continue
            v = np.asarray(v, dtype=float)
            if not np.all(v == v[..., :1]):
                raise NotImplementedError("re-discretisation needs one bond length per chain")
            out[k] = np.repeat(v[..., :1], n_bonds, axis=-1)
        return out

    def _like(self, r, t3, t2, states, mods, grid):
        N = r.shape[1]
        spacing = 16.5 if self.min_spacing is None else self.min_spacing
        g = dict(grid, vf_limit=self.grid.get("vf_limit", 0.5))
        return ReplicaEnsemble(r, t3, t2, states, mods, binders=self.binders, bond_params=self._uniform_bonds(N - 1),
                               grid=g, bead_vol=self.bead_vol, chi=self.chi, mu=self.mu, max_binders=self.max_binders,
                               moves=default_moves(self.R, N, spacing), min_spacing=self.min_spacing,
                               device=self.device, assume_fully_accessible=self.assume_fully_accessible,
                               replica_offset=self.replica_offset, fast_field_points=self.fast_field_points)

    def coarse_grained(self, cg_factor: int):
        """Every replica coarse-grained by `cg_factor` (get_cg_chromatin + get_cg_udf, rediscretize.py:
        199-272, 401-471) in one launch; returns a new ensemble on the coarser grid."""
        from .util import rediscretize as rd
        self.engine.sync()
        self.pull()
        cg = rd.coarse_grain_ensemble(self.r, self.t3, self.states, self.chemical_mods, cg_factor, self.device)
        return self._like(cg["r"], cg["t3"], cg["t2"], cg["states"], cg["chemical_mods"],
                          rd.cg_grid(self.grid, cg_factor))

    def refined(self, num_beads_refined: int, bead_spacing: float, chemical_mods, seed: Optional[int] = 0,
                binding_equilibration: int = 0):
        """Every replica refined to `num_beads_refined` beads (refine_chromatin + refine_udf,
        rediscretize.py:845-900, 994-1107): Brownian bridges between the coarse beads, scaled outwards,
        pulled back into a spherical confinement, states reset to 0 and -- optionally -- equilibrated by
        `binding_equilibration` binding-state attempts per replica in one launch of the MC kernel."""
        from .util import rediscretize as rd
        self.engine.sync()
        self.pull()
        M, g = self.N, self.grid
        if rd.refined_num_points(M, num_beads_refined) != num_beads_refined:
            raise ValueError("num_beads_refined must not be a multiple of (coarse beads - 1): the reference's "
                             "path then has one bead fewer (rediscretize.py:571)")
        fine_grid = rd.refined_grid(dict(nx=g["nx"], ny=g["ny"], nz=g["nz"], dx=g["x_width"] / g["nx"],
                                         dy=g["y_width"] / g["ny"], dz=g["z_width"] / g["nz"],
                                         confine_type=g["confine_type"], confine_length=g["confine_length"]),
                                    num_beads_refined / M)
        rad = fine_grid["confine_length"] if fine_grid["confine_type"] == "Spherical" else 0.0
        fine = rd.refine_ensemble(self.r, self.t3, num_beads_refined, bead_spacing, confine_length=rad, seed=seed,
                                  device=self.device)
        mods = np.ascontiguousarray(np.broadcast_to(np.asarray(chemical_mods, dtype=np.int64).reshape(
            (-1, num_beads_refined, self.nb)), (self.R, num_beads_refined, self.nb)))
        ens = self._like(fine["r"], fine["t3"], fine["t2"], np.zeros_like(mods), mods, fine_grid)
        if binding_equilibration > 0:  # rediscretize.py:1093-1106: binding-only attempts, one bead each
            keep = ens.moves.copy()
            mv = keep.copy()
            mv["move_on"][:] = 0
            mv["move_on"][:, 4], mv["num_per_cycle"][:, 4], mv["controller"][:, 4] = 1, binding_equilibration, 0
            ens.moves = mv
            ens.engine.set_moves(mv)
            ens.mc_sim(1, 1.0, 0 if seed is None else seed, sync_host=False)
            ens.engine.sync()
            ens.pull()
            ens.moves = keep
            ens.engine.set_moves(keep)
        return ens

    def close(self):
        if self._pinned:  # before the arrays can be freed
            from .engine import host_unregister
            self.engine.sync()
            for a in self._pinned:
                host_unregister(a)
            self._pinned = []
        self.engine.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- construction helpers ----------------------------------------------
    @classmethod
    def from_polymers(cls, polymers, fields, controllers=None, device: int = 0):
        """Stack reference-style objects (one polymer + one field per replica)."""
        from .fields import binder_dicts
        from .mc.mc_sim import controllers_to_moves
        p0, f0 = polymers[0], fields[0]
        st = lambda name: np.stack([getattr(p, name) for p in polymers])
        grid = f0._grid()
        bd = binder_dicts(p0)
        # the field's own binder table (fields.pyx:687-712 reads it from the DataFrame snapshot) carries
        # what a re-discretised ensemble needs to rebuild its prefactors on another grid
        for b, rec in zip(bd, getattr(f0, "binder_dict", None) or []):
            b["interaction_energy"] = float(rec["interaction_energy"])
            b["interaction_volume"] = float(rec["interaction_volume"])
            b["interaction_radius"] = float((3.0 * rec["interaction_volume"] / (4.0 * np.pi)) ** (1.0 / 3.0))
            b["cross_talk"] = dict(rec.get("cross_talk_interaction_energy", {}) or {})
        pre = f0._prefactors(p0.num_binders)
        keys = ("eps_bend", "eps_par", "eps_perp", "gamma", "eta")
        if getattr(p0, "eps_twist", None) is not None:  # SSTWLC replicas keep their twist term
            keys += ("eps_twist", "natural_twist")
        bond = {k: np.stack([getattr(p, k) for p in polymers]) for k in keys}
        if all(np.array_equal(bond["eps_bend"][0], b) for b in bond["eps_bend"]):
            bond = {k: v[0] for k, v in bond.items()}
        if getattr(p0, "nucleosome_constants", None) is not None:  # DetailedChromatin replicas share one bp_wrap
            bond["nucleosome_constants"] = np.asarray(p0.nucleosome_constants, dtype=float)
        moves = None
        if controllers is not None:
            moves = np.stack([controllers_to_moves(c) for c in controllers])
        mu = np.array([[b["chemical_potential"] for b in binder_dicts(p)] for p in polymers])
        ens = cls(st("r"), st("t3"), st("t2"), st("states"), st("chemical_mods"), binders=bd, bond_params=bond,
                  grid=grid, bead_vol=p0.beads[0].vol, chi=[getattr(f, "chi", 1.0) for f in fields], mu=mu,
                  max_binders=p0.max_binders, moves=moves, min_spacing=float(np.min(p0.bead_length)),
                  access_vol=None if getattr(f0, "assume_fully_accessible", 1) == 1 else f0.access_vols,
                  device=device, field_prefactors=pre,
                  assume_fully_accessible=getattr(f0, "assume_fully_accessible", 1),
                  fast_field_points=int(getattr(f0, "n_points", 0)) if getattr(f0, "fast_field", 0) == 1 else 0)
        if controllers is not None:  # every replica goes through the move types in its controller list's order
            from .mc.mc_sim import controller_order
            orders = {tuple(controller_order(c)) for c in controllers}
            if len(orders) != 1:
                raise ValueError("the replicas of one ensemble must list their move controllers in the same order")
            ens.engine.set_move_order(orders.pop())
        return ens
