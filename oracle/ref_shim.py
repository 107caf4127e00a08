"""Import the UNMODIFIED reference module from /root/reference (test infrastructure only).

The reference package's ``__init__`` eagerly imports matplotlib, OpenEXR and torchmetrics,
none of which exist in this image.  The shim registers an empty package object whose
``__path__`` points at the reference tree (so the eager ``__init__`` never runs), stubs
``matplotlib.pyplot`` and imports just the two modules on the hot path.  It is loaded under
an alias package name so it can live beside the drop-in module in one process.

``/root/reference`` does not exist on the GPU box; there the offline install under
``baseline/_ref`` (same files, unmodified) is used.  Callers must check ``available()``.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

_ALIAS = "_lhg_reference_pkg"
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    """LHG_REFERENCE_ROOT, else the read-only checkout, else the offline install of the unmodified reference under
    baseline/_ref (git-ignored; made by __graft_entry__.build() with pip --target, it travels to the GPU box)."""
    for root in (os.environ.get("LHG_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if root and os.path.isfile(os.path.join(root, "learnedMethodForHologram", "angular_spectrum_method.py")):
            return root
    return os.environ.get("LHG_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isfile(
        os.path.join(REFERENCE_ROOT, "learnedMethodForHologram", "angular_spectrum_method.py")
    )


def load():
    """Return (angular_spectrum_method, utilities) modules of the reference."""
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    if _ALIAS + ".angular_spectrum_method" in sys.modules:
        return (
            sys.modules[_ALIAS + ".angular_spectrum_method"],
            sys.modules[_ALIAS + ".utilities"],
        )
    if "matplotlib" not in sys.modules:
        try:
            importlib.import_module("matplotlib.pyplot")
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    pkg = types.ModuleType(_ALIAS)
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "learnedMethodForHologram")]
    sys.modules[_ALIAS] = pkg
    util = importlib.import_module(_ALIAS + ".utilities")
    asm = importlib.import_module(_ALIAS + ".angular_spectrum_method")
    return asm, util


def load_next():
    """The reference modules either side of the path: returns a dict with ``loss_func``,
    ``data_loader``, ``neural_network_components``, ``AP2POH`` and ``utilities``.  The sub-package's eager
    ``__init__`` (which pulls in torchmetrics through watermelon.py) is bypassed the same way as the parent's."""
    load()
    sub = _ALIAS + ".watermelon_hologram"
    if sub not in sys.modules:
        pkg = types.ModuleType(sub)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "learnedMethodForHologram", "watermelon_hologram")]
        sys.modules[sub] = pkg
    return {
        "utilities": sys.modules[_ALIAS + ".utilities"],
        "neural_network_components": importlib.import_module(_ALIAS + ".neural_network_components"),
        "loss_func": importlib.import_module(sub + ".loss_func"),
        "data_loader": importlib.import_module(sub + ".data_loader"),
        "AP2POH": importlib.import_module(sub + ".AP2POH"),
    }
