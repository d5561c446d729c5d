"""Monte-Carlo driver (host mirror of chromo/mc/__init__.py).

`polymer_in_field` keeps the reference's snapshot loop (mc/__init__.py:36-160):
per snapshot an annealing factor, an inner seed, one `mc_sim` call, then the
configuration is written as a reference-schema CSV (and handed to `save_snapshot`);
`continue_polymer_in_field_simulation` resumes from the latest one.  The parameter
logging of `make_reproducible` is out of scope; its folder layout is kept.
"""
from pathlib import Path
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .mc_controller import Controller, SimpleControl, all_moves
from .mc_sim import mc_sim, mc_step, set_rng_mode, rng_mode  # noqa: F401
from .moves import Bounds
from ..util.poly_stat import find_polymers_in_output_dir, get_latest_configuration, get_latest_simulation
from ..util.reproducibility import get_unique_subfolder, sim_folder_prefix


def get_amplitude_bounds(polymers) -> Tuple[Bounds, Bounds]:
    """Lower / upper bounds of the bead-selection and move amplitudes
    (mc/__init__.py:295-332)."""
    poly_len = np.min([polymer.r.shape[0] for polymer in polymers])
    min_spacing = np.min([np.min(polymer.bead_length) for polymer in polymers])
    bead_amp_bounds = Bounds("bead_amp_bounds", {
        "crank_shaft": (min(30, poly_len), min(150, poly_len)),
        "slide": (min(10, poly_len), min(150, poly_len)),
        "end_pivot": (min(50, poly_len / 4), min(150, int(poly_len / 2))),
        "tangent_rotation": (1, poly_len),
        "change_binding_state": (1, 1),
    })
    move_amp_bounds = Bounds("move_amp_bounds", {
        "crank_shaft": (0.1 * np.pi, 0.25 * np.pi),
        "slide": (0.2 * min_spacing, 0.3 * min_spacing),
        "end_pivot": (0.2 * np.pi, 0.25 * np.pi),
        "tangent_rotation": (0.05 * np.pi, 0.2 * np.pi),
        "change_binding_state": (0, 0),
    })
    return bead_amp_bounds, move_amp_bounds


def polymer_in_field(polymers, binders, field, num_save_mc, num_saves, bead_amp_bounds, move_amp_bounds,
                     mc_move_controllers: Optional[List[Controller]] = None, random_seed: Optional[int] = 0,
                     mu_schedule=None, output_dir: Optional[str] = None,
                     save_snapshot: Optional[Callable] = None, continue_from: Optional[str] = None, **kwargs):
    """Monte-Carlo simulation of polymers in a field, `num_saves` snapshots of
    `num_save_mc` sweeps each (mc/__init__.py:36-160).

    With `output_dir`, the run gets its own folder `output_dir/sim_<k>/` like a
    `make_reproducible` run of the reference: the initial configuration of every polymer
    (file `<name>`), `<name>-<snapshot>.csv` after every snapshot and the per-move
    acceptance logs under `acceptance_trackers/`.  `continue_from` names the folder of the
    run being continued: snapshot numbering goes on where it stopped.  Without `output_dir`
    nothing is written (`save_snapshot` still sees every snapshot).  Returns the polymers.
    """
    np.random.seed(random_seed)
    run_dir = None
    first_snapshot = 0
    if output_dir is not None:
        run_dir = get_unique_subfolder(Path(output_dir) / sim_folder_prefix)
        for poly in polymers:
            poly.to_file(str(run_dir / poly.name))
        if continue_from is not None:
            prev = Path(output_dir) / continue_from
            done = [int(f.stem.split("-")[-1]) for f in prev.glob("*-*.csv") if f.stem.split("-")[-1].isdigit()]
            first_snapshot = max(done) + 1 if done else 0
    if mc_move_controllers is None:
        mc_move_controllers = all_moves(log_dir=str(run_dir) if run_dir is not None else ".",
                                        bead_amp_bounds=bead_amp_bounds.bounds,
                                        move_amp_bounds=move_amp_bounds.bounds, controller=SimpleControl)
    elif run_dir is not None:
        for controller in mc_move_controllers:
            controller.move.acceptance_tracker.log_dir = str(run_dir) + '/acceptance_trackers'
    for k in range(num_saves):
        mc_count = first_snapshot + k
        if mu_schedule is not None:
            mu_adjust_factor = mu_schedule.function(k, num_saves)
        else:
            mu_adjust_factor = 1
        inner_seed = np.random.randint(0, 1E9)
        mc_sim(polymers, binders, num_save_mc, mc_move_controllers, field, mu_adjust_factor, inner_seed)
        if run_dir is not None:
            for poly in polymers:
                poly.to_csv(str(run_dir / f"{poly.name}-{mc_count}.csv"))
            for controller in mc_move_controllers:
                controller.move.acceptance_tracker.create_log_file(mc_count)
                controller.move.acceptance_tracker.save_move_log(snapshot=mc_count)
        if save_snapshot is not None:
            save_snapshot(mc_count, polymers, field, mc_move_controllers)
    if run_dir is not None:
        for poly in polymers:
            poly.update_log_path(f"{run_dir}/{poly.name}_config_log.csv")
    return polymers


_polymer_in_field = polymer_in_field


def continue_polymer_in_field_simulation(polymer_class, binders, field, output_dir: str, num_save_mc: int,
                                         num_saves: int, mc_move_controllers: Optional[List[Controller]] = None,
                                         random_seed: Optional[int] = 0):
    """Continue the latest run found under `output_dir` (mc/__init__.py:167-227): load
    every polymer's latest snapshot, attach the polymers to `field` (densities are
    recomputed from the loaded configuration), and run `num_saves` more snapshots into a
    new `sim_<k>` folder.  Returns the polymers."""
    latest = get_latest_simulation(output_dir)
    latest_path = f"{output_dir}/{latest}"
    names = find_polymers_in_output_dir(latest_path)
    if not names:
        raise FileNotFoundError(f"no polymer configurations (files 'Chr-<i>') in {latest_path}")
    paths = [get_latest_configuration(polymer_prefix=nm, directory=latest_path) for nm in names]
    labels = ["-".join(p.split("/")[-1].split(".")[0].split("-")[0:2]) for p in paths]
    polymers = [polymer_class.from_file(paths[i], labels[i]) for i in range(len(paths))]
    field.polymers = polymers
    if hasattr(field, "update_all_densities_for_all_polymers"):
        field.update_all_densities_for_all_polymers()
    bead_amp_bounds, move_amp_bounds = get_amplitude_bounds(polymers)
    return polymer_in_field(polymers, binders, field, num_save_mc, num_saves, bead_amp_bounds, move_amp_bounds,
                            mc_move_controllers=mc_move_controllers, random_seed=random_seed,
                            output_dir=output_dir, continue_from=latest)
