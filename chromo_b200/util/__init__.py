"""Host-side utilities.  `dss_params` is the dssWLC parameter table the
reference loads from chromo/util/dssWLCparams (chromo/util/__init__.py:5-6);
columns: delta, eps_bend, gamma, eps_par, eps_perp, eta (Koslover & Spakowitz,
Soft Matter 2013)."""
from pathlib import Path

import numpy as np

dss_params = np.load(Path(__file__).resolve().parents[1] / "data" / "dsswlc_params.npy")
