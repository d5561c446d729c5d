"""Nucleosome geometry constants for `DetailedChromatin` (host mirror of chromo/util/nucleo_geom.py:17-245 and
DetailedNucleosome.__init__, chromo/beads.py:448-515).

A detailed nucleosome wraps `bp_wrap` base pairs of DNA on a left-handed super-helix of radius R and pitch h.  In
its own frame the entering DNA has tangent `t3_local` and normal `t2_local`, enters at `r_enter_local` and leaves
at `r_exit_local`; the exiting tangent / binormal are fixed linear combinations (`a3`, `a1`) of the entering
(t3, t2, t1).  The kernels take these 20 numbers (`chromo_set_detailed_nucleosomes`) and build the per-bead
rotation themselves."""
import numpy as np

NUC_TWIST_DENS = 10.17
LENGTH_BP = 0.332
R_default = 4.1899999999999995
h_default = 4.531142964071856 / 2
bp_wrap_default = 147
s_default = (bp_wrap_default - 1) * LENGTH_BP
w0_default = 2 * np.pi / (NUC_TWIST_DENS * LENGTH_BP)
Lt_default = np.sqrt(4 * np.pi ** 2 * R_default ** 2 + h_default ** 2)
Phi_default = w0_default - 2 * np.pi * h_default / (Lt_default ** 2)
consts_dict = {"R": R_default, "h": h_default, "bp_wrap": bp_wrap_default, "length_bp": LENGTH_BP, "s": s_default,
               "nuc_twist_dens": NUC_TWIST_DENS, "w0": w0_default, "Lt": Lt_default, "Phi": Phi_default}


def t3(s):
    R, Lt, h = R_default, Lt_default, h_default
    return np.array([-2 * np.pi * R / Lt * np.sin(2 * np.pi * s / Lt), 2 * np.pi * R / Lt * np.cos(2 * np.pi * s / Lt),
                     h / Lt])


def normal(s):
    Lt = Lt_default
    return np.array([-np.cos(2 * np.pi * s / Lt), -np.sin(2 * np.pi * s / Lt), 0])


def binormal(s):
    return np.cross(t3(s), normal(s))


def t1(s):
    return np.cos(Phi_default * s) * normal(s) + np.sin(Phi_default * s) * binormal(s)


def t2(s):
    return -np.sin(Phi_default * s) * normal(s) + np.cos(Phi_default * s) * binormal(s)


def get_r(s):
    """Position of the DNA path at wrapped length s in the nucleosome's frame (nucleo_geom.py:275-294)."""
    R, Lt, h = R_default, Lt_default, h_default
    return np.array([R * np.cos(2 * np.pi * s / Lt), R * np.sin(2 * np.pi * s / Lt), h * s / Lt - ((s_default / Lt * h) / 2)])


def nucleosome_constants(bp_wrap: float, with_diameter: bool = True) -> np.ndarray:
    """The 20 constants of one `bp_wrap`, in the order `chromo_set_detailed_nucleosomes` takes them.
    `with_diameter=False` (DetailedChromatin2, polymers.pyx:2627-2735): entry / exit offsets of length zero -- the
    bonds run between the bead centres and only the exit frame enters the energy."""
    s = (bp_wrap - 1) * LENGTH_BP
    R, Lt, h = R_default, Lt_default, h_default
    t3_local = np.array([0, 2 * np.pi * R / Lt, h / Lt])
    t2_local = np.array([0, -h / Lt, 2 * np.pi * R / Lt])
    r_enter = np.array([R, 0, -(h * s_default / Lt) / 2])
    r_exit = get_r(s)
    ne, nx = np.linalg.norm(r_enter), np.linalg.norm(r_exit)
    r_enter, r_exit = r_enter / ne, r_exit / nx
    if not with_diameter:
        ne = nx = 0.0
    a3 = [np.dot(t3(s), t3(0)), np.dot(t3(s), t2(0)), np.dot(t3(s), t1(0))]
    a1 = [np.dot(t1(s), t3(0)), np.dot(t1(s), t2(0)), np.dot(t1(s), t1(0))]
    return np.concatenate([t3_local, t2_local, r_enter, [ne], r_exit, [nx], a3, a1]).astype(float)
