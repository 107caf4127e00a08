"""B200-native band-limited angular-spectrum propagation (drop-in for the hot path of
WeijieXie/learned_hologram_gan).  Importing the package does not need a GPU; constructing a
propagator does, and nothing here falls back to a CPU or PyTorch implementation."""

from . import _cabi  # noqa: F401  (fails loudly if the library cannot be described)
from .angular_spectrum_method import (  # noqa: F401
    bandLimitedAngularSpectrumMethod,
    bandLimitedAngularSpectrumMethod_for_multiple_distances,
    bandLimitedAngularSpectrumMethod_for_single_fixed_distance,
)

__all__ = [
    "bandLimitedAngularSpectrumMethod",
    "bandLimitedAngularSpectrumMethod_for_single_fixed_distance",
    "bandLimitedAngularSpectrumMethod_for_multiple_distances",
]
