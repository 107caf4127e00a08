"""Import shim: with this repository on ``sys.path`` the reference's import lines keep working
and ``learnedMethodForHologram.angular_spectrum_method`` is the B200-native module.

Set ``LHG_REFERENCE_ROOT`` to a checkout of WeijieXie/learned_hologram_gan to serve its other
submodules (networks, training loop, data tooling) unchanged next to the replaced module."""

import sys as _sys

from learned_hologram_gan_b200 import overlay as _overlay

_overlay.install(package=_sys.modules[__name__])
