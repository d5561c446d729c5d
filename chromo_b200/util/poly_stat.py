"""Locating the configurations of an earlier run (the three helpers of
chromo/util/poly_stat.py:184-261 that `continue_polymer_in_field_simulation` uses; the
polymer statistics of that module are out of scope)."""
import os
from typing import List, Optional


def get_latest_simulation(directory: Optional[str] = '.') -> str:
    """Name of the sub-folder `<prefix>_<k>` with the largest k."""
    sims = [d for d in os.listdir(directory) if os.path.isdir(os.path.join(directory, d))]
    if not sims:
        raise FileNotFoundError(f"no simulation folders in {directory}")
    return max(sims, key=lambda d: int(d.split("_")[-1]))


def find_polymers_in_output_dir(directory: Optional[str] = '.') -> List[str]:
    """Polymer names `Chr-<i>` that have an initial-configuration file (no .csv suffix) in
    `directory`, sorted."""
    ids = {f.split("-")[1] for f in os.listdir(directory)
           if os.path.isfile(os.path.join(directory, f)) and f.startswith("Chr-") and not f.endswith(".csv")}
    return ["Chr-" + i for i in sorted(ids)]


def get_latest_configuration(polymer_prefix: Optional[str] = "Chr-1", directory: Optional[str] = '.') -> str:
    """Path of `<polymer_prefix>-<k>.csv` with the largest snapshot index k."""
    snaps = [int(f.split("-")[2].split(".")[0]) for f in os.listdir(directory)
             if f.startswith(polymer_prefix + "-") and f.endswith(".csv")]
    if not snaps:
        raise FileNotFoundError(f"no snapshots of {polymer_prefix} in {directory}")
    return f"{directory}/{polymer_prefix}-{max(snaps)}.csv"
