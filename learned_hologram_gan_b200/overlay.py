"""Make ``learnedMethodForHologram.angular_spectrum_method`` resolve to the B200 path.

The reference's callers import the propagators as
``from learnedMethodForHologram.angular_spectrum_method import ...`` (generatePOH.py:8-10,
generator.py:10-12, AP2POH.py:8-10, watermelon.py:19-21) and its test reaches the module as an
attribute of the package after importing only ``learnedMethodForHologram.utilities``
(tests/test_angular_spectrum_method.py:3,16).  ``install`` registers a package object of that
name whose ``angular_spectrum_method`` is this repo's module; every other submodule is served
from the reference tree when one is given (its files stay untouched), and ``utilities`` falls
back to this repo's subset when the reference tree (or matplotlib) is absent.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

PACKAGE = "learnedMethodForHologram"


def install(reference_root: str | None = None, package: types.ModuleType | None = None):
    from . import angular_spectrum_method as asm_mod
    from . import utilities as util_mod

    reference_root = reference_root or os.environ.get("LHG_REFERENCE_ROOT")
    pkg = package or sys.modules.get(PACKAGE)
    if pkg is None:
        pkg = types.ModuleType(PACKAGE)
        pkg.__path__ = []
        sys.modules[PACKAGE] = pkg
    ref_dir = os.path.join(reference_root, PACKAGE) if reference_root else None
    if ref_dir and os.path.isdir(ref_dir) and ref_dir not in list(pkg.__path__):
        pkg.__path__.append(ref_dir)
    # the propagation module is always ours
    sys.modules[PACKAGE + ".angular_spectrum_method"] = asm_mod
    pkg.angular_spectrum_method = asm_mod
    # utilities: the reference's full module if it imports here, else our subset
    util = None
    if ref_dir and os.path.isdir(ref_dir):
        try:
            util = importlib.import_module(PACKAGE + ".utilities")
        except Exception:
            sys.modules.pop(PACKAGE + ".utilities", None)
            util = None
    if util is None:
        util = util_mod
        sys.modules[PACKAGE + ".utilities"] = util
    pkg.utilities = util
    # The reference's sub-package __init__ imports every module it has, watermelon.py (torchmetrics) included.  A
    # package object that only carries the directory lets `from learnedMethodForHologram.watermelon_hologram.X
    # import ...` load exactly the module asked for (generatePOH.py:4-7, trainingModel.py), each of them unmodified.
    sub = PACKAGE + ".watermelon_hologram"
    sub_dir = os.path.join(ref_dir, "watermelon_hologram") if ref_dir else None
    if sub_dir and os.path.isdir(sub_dir) and sub not in sys.modules:
        spkg = types.ModuleType(sub)
        spkg.__path__ = [sub_dir]
        spkg.__package__ = sub
        sys.modules[sub] = spkg
        pkg.watermelon_hologram = spkg
    return pkg
