"""Reader proteins (binders) -- host-side parameter objects.

Mirrors chromo/binders.pyx (Binder 16-49, ReaderProtein 51-131, the module-level
singletons null_reader / hp1 / prc1 133-173, get_by_name 185-207,
make_binder_collection 210-238).  Nothing here runs on the hot path: the values
are packed into the device tables by chromo_b200.fields / chromo_b200.ensemble.
"""
import inspect
import sys

import numpy as np
import pandas as pd


class Binder:
    """An arbitrary component binding to a polymer (binders.pyx:16-49)."""

    def __init__(self, name: str, sites_per_bead: int) -> None:
        self.name = name
        self.sites_per_bead = sites_per_bead
        self.binding_seq = np.zeros((sites_per_bead,), dtype=int)


class ReaderProtein(Binder):
    """Chemical properties of a reader protein (binders.pyx:51-131)."""

    def __init__(self, name, sites_per_bead, bind_energy_mod, bind_energy_no_mod, interaction_energy,
                 chemical_potential, interaction_radius, cross_talk_interaction_energy=None) -> None:
        super().__init__(name, sites_per_bead)
        self.bind_energy_mod = bind_energy_mod
        self.bind_energy_no_mod = bind_energy_no_mod
        self.interaction_energy = interaction_energy
        self.chemical_potential = chemical_potential
        self.interaction_radius = interaction_radius
        self.interaction_volume = (4.0 / 3.0) * np.pi * interaction_radius ** 3
        self.field_energy_prefactor = 0.0
        self.interaction_energy_intranucleosome = 0.0
        self.cross_talk_interaction_energy = (
            {} if cross_talk_interaction_energy is None else cross_talk_interaction_energy)
        self.cross_talk_field_energy_prefactor = {}

    def dict(self):
        return {
            "name": self.name,
            "sites_per_bead": self.sites_per_bead,
            "bind_energy_mod": self.bind_energy_mod,
            "bind_energy_no_mod": self.bind_energy_no_mod,
            "interaction_energy": self.interaction_energy,
            "chemical_potential": self.chemical_potential,
            "interaction_radius": self.interaction_radius,
            "interaction_volume": self.interaction_volume,
            "field_energy_prefactor": self.field_energy_prefactor,
            "interaction_energy_intranucleosome": self.interaction_energy_intranucleosome,
            "cross_talk_interaction_energy": self.cross_talk_interaction_energy,
            "cross_talk_field_energy_prefactor": self.cross_talk_field_energy_prefactor,
        }


null_reader = ReaderProtein('null_reader', sites_per_bead=0, bind_energy_mod=0, bind_energy_no_mod=0,
                            interaction_energy=0, chemical_potential=0, interaction_radius=0)
"""Placeholder: the simulator needs at least one reader protein (binders.pyx:133-143)."""

hp1 = ReaderProtein('HP1', sites_per_bead=2, bind_energy_mod=-0.01, bind_energy_no_mod=1.52,
                    interaction_energy=-4, chemical_potential=-1, interaction_radius=3,
                    cross_talk_interaction_energy={'PRC1': 0})
"""Heterochromatin Protein 1, binds H3K9me marks (binders.pyx:146-166)."""

prc1 = ReaderProtein('PRC1', sites_per_bead=2, bind_energy_mod=-0.01, bind_energy_no_mod=1.52,
                     interaction_energy=-4, chemical_potential=-1, interaction_radius=3,
                     cross_talk_interaction_energy={'HP1': 0})
"""Polycomb repressive complex 1, binds H3K27me marks (binders.pyx:169-182)."""


def get_by_name(name):
    """Look up a saved reader protein by name (binders.pyx:185-207)."""
    all_binders = [obj for _, obj in inspect.getmembers(sys.modules[__name__]) if isinstance(obj, Binder)]
    matching = [b for b in all_binders if b.name == name]
    if not matching:
        raise ValueError(f"No binders found in {__name__} with name: {name}")
    if len(matching) > 1:
        raise ValueError(f"More than one binder has the name requested: {name}")
    return matching[0]


def make_binder_collection(binders):
    """Summary DataFrame of a sequence of binders (binders.pyx:210-238)."""
    df = pd.DataFrame(columns=['name', 'sites_per_bead', 'bind_energy', 'interaction_energy',
                               'chemical_potential'])
    if binders is None:
        return None
    if type(binders) is str or issubclass(type(binders), Binder):
        binders = [binders]
    rows = []
    for binder in binders:
        if type(binder) is str:
            binder = get_by_name(binder)
        rows.append(binder.dict())
    if rows:
        df = pd.concat([df, pd.DataFrame(rows)], ignore_index=True)
    return df
