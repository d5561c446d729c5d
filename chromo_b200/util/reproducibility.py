"""Run-directory layout of a simulation (the part of chromo/util/reproducibility.py the
snapshot / resume path needs; the parameter-logging decorator itself is out of scope).

    output_dir/
        sim_1/  sim_2/ ...            one folder per `polymer_in_field` call
            acceptance_trackers/      per-move acceptance logs
            <polymer name>            the initial configuration (what find_polymers_in_output_dir looks for)
            <polymer name>-<k>.csv    configuration after snapshot k
"""
from pathlib import Path

sim_folder_prefix = "sim_"


def get_unique_subfolder(root) -> Path:
    """Create `<root>1`, `<root>2`, ... -- the first that does not exist yet -- with its
    `acceptance_trackers/` folder, and return it (reproducibility.py:325-359; relies on
    mkdir being atomic, so concurrent runs get different folders)."""
    i = 1
    while True:
        folder = Path(f"{root}{i}")
        try:
            folder.mkdir(parents=True)
        except FileExistsError:
            i += 1
            continue
        (folder / "acceptance_trackers").mkdir()
        return folder


def get_unique_subfolder_name(root) -> Path:
    """The folder `get_unique_subfolder` would create next (reproducibility.py:362-382)."""
    i = 1
    while Path(f"{root}{i}").is_dir():
        i += 1
    return Path(f"{root}{i}")
