"""ctypes binding of include/lhg_next_b200.h (the stages either side of the propagation path).

Same library as ``_cabi`` (``libasm_b200.so``), same rules: raw device pointers and sizes only, the stream as
``torch.cuda.current_stream().cuda_stream``, no fallback when the library is missing.
"""

from __future__ import annotations

import ctypes as C

from . import _cabi

STATUS_NAMES = {0: "LHG_OK", -1: "LHG_EINVAL", -3: "LHG_ECUDA", -4: "LHG_EWORKSPACE"}

EXPORTS = (
    "lhg_next_version",
    "lhg_next_last_error",
    "lhg_next_launch_count",
    "lhg_next_partial_floats",
    "lhg_amp_loss_terms",
    "lhg_amp_loss_backward",
    "lhg_focal_phase_loss_terms",
    "lhg_focal_phase_loss_backward",
    "lhg_phase_gradient_loss_backward",
    "lhg_phase_point_loss_terms",
    "lhg_phase_point_loss_backward",
    "lhg_plane_minmax",
    "lhg_normalize_planes",
    "lhg_amplitude_normalize",
    "lhg_pack_rgb_u8",
    "lhg_ap2poh_tail",
    "lhg_ap2poh_tail_backward_floats",
    "lhg_ap2poh_tail_backward",
    "lhg_bin_gather",
    "lhg_assemble_rgbd",
    "lhg_scale_two_pi",
)


class NextError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


_bound = False


def load():
    """The shared library with the argument types of this header declared (once)."""
    global _bound
    lib = _cabi.load()
    if _bound:
        return lib
    P, LL, I, F, SZ = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_size_t
    lib.lhg_next_version.restype = I
    lib.lhg_next_last_error.restype = C.c_char_p
    lib.lhg_next_launch_count.restype = LL
    lib.lhg_next_partial_floats.restype = SZ
    lib.lhg_next_partial_floats.argtypes = [LL, I, I]
    lib.lhg_ap2poh_tail_backward_floats.restype = SZ
    lib.lhg_ap2poh_tail_backward_floats.argtypes = [I, LL, I, I]
    sigs = {
        "lhg_amp_loss_terms": [P, P, LL, I, I, F, P, SZ, P, P],
        "lhg_amp_loss_backward": [P, P, P, P, F, LL, I, I, P, P],
        "lhg_focal_phase_loss_terms": [P, P, LL, I, I, P, SZ, P, P],
        "lhg_focal_phase_loss_backward": [P, P, P, P, LL, I, I, P, P],
        "lhg_phase_gradient_loss_backward": [P, P, P, LL, I, I, P, P],
        "lhg_phase_point_loss_terms": [P, P, LL, I, I, P, SZ, P, P],
        "lhg_phase_point_loss_backward": [P, P, P, P, I, LL, I, I, P, P],
        "lhg_plane_minmax": [P, LL, LL, P, SZ, P, P],
        "lhg_normalize_planes": [P, P, LL, LL, P, P],
        "lhg_amplitude_normalize": [P, P, LL, LL, P, P],
        "lhg_pack_rgb_u8": [P, P, LL, I, I, I, P, P],
        "lhg_ap2poh_tail": [P, P, P, I, LL, I, I, P, SZ, P, P, P],
        "lhg_ap2poh_tail_backward": [P, P, P, I, P, LL, I, I, P, SZ, P, P, P, P],
        "lhg_bin_gather": [P, LL, SZ, SZ, P, I, P, I],
        "lhg_assemble_rgbd": [P, P, I, LL, LL, P, P],
        "lhg_scale_two_pi": [P, LL, P, P],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.restype = I
        fn.argtypes = argtypes
    _bound = True
    return lib


def check(code):
    if code != 0:
        raise NextError(code, load().lhg_next_last_error().decode("utf-8", "replace"))
