"""chromo_b200 -- B200-native Monte-Carlo energy-evaluation path of chromo.

Keeps the reference's Python API surface for the hot path
(`polymers.Chromatin/SSWLC`, `binders`, `fields.UniformDensityField`, the `mc`
move classes and the `mc_sim` driver) on top of hand-written sm_100a CUDA
kernels reached through a C ABI (include/chromo_b200.h).
"""
__version__ = "0.1.0"
