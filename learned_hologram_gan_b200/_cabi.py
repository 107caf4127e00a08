"""ctypes binding of include/asm_b200.h (the drop-in boundary).

No torch types cross this boundary: tensors are passed as raw device pointers plus sizes,
the stream as ``torch.cuda.current_stream().cuda_stream``.  There is NO fallback: if the
shared library is missing the import of the product path raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LHG_LIB") or os.path.join(_HERE, "lib", "libasm_b200.so")

# enums (keep in sync with include/asm_b200.h)
IN_PHASE, IN_AMP_PHASE, IN_COMPLEX, IN_SPECTRUM, IN_COTANGENT = range(5)
FILTER_NONE, FILTER_H = 0, 1
FLAG_CONJ, FLAG_CIRC_MASK = 1, 2
OUT_ABS, OUT_ANGLE, OUT_ABS_ANGLE, OUT_COMPLEX, OUT_ABS2, OUT_SPECTRUM, OUT_GRAD_PHASE = range(7)
GRID_W, GRID_CIRC_MASK, GRID_RADIAL, GRID_H, GRID_BAND_LIMIT = range(5)

STATUS_NAMES = {0: "ASM_OK", -1: "ASM_EINVAL", -2: "ASM_EUNSUPPORTED_SIZE", -3: "ASM_ECUDA", -4: "ASM_EWORKSPACE"}

EXPORTS = (
    "asm_version",
    "asm_sizeof_io",
    "asm_last_error",
    "asm_plan_create",
    "asm_plan_destroy",
    "asm_plan_info",
    "asm_workspace_bytes",
    "asm_build_grid",
    "asm_wm_tiled_bytes",
    "asm_build_wm_tiled",
    "asm_propagate",
    "asm_launch_count",
    "asm_loss_finish",
    "asm_fused_step_supported",
    "asm_profile_enable",
    "asm_profile_collect",
)


class AsmIO(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32),
        ("n_samples", C.c_int32),
        ("n_depth", C.c_int32),
        ("reduce_depth", C.c_int32),
        ("in_kind", C.c_int32),
        ("filter_kind", C.c_int32),
        ("filter_flags", C.c_int32),
        ("out_kind", C.c_int32),
        ("in0", C.c_void_p),
        ("in1", C.c_void_p),
        ("cot_abs", C.c_void_p),
        ("cot_angle", C.c_void_p),
        ("cot_abs2", C.c_void_p),
        ("cot_target", C.c_void_p),
        ("cot_scale", C.c_float),
        ("phase_scale", C.c_float),
        ("wm_grid", C.c_void_p),
        ("z_dev", C.c_void_p),
        ("depth_index", C.c_void_p),
        ("n_z", C.c_int32),
        ("reserved0", C.c_int32),
        ("out0", C.c_void_p),
        ("out1", C.c_void_p),
        ("save_field", C.c_void_p),
        ("aux_phase", C.c_void_p),
        ("aux_amp", C.c_void_p),
        ("out_scale", C.c_float),
        ("loss_partial_len", C.c_int32),
        ("loss_target", C.c_void_p),
        ("loss_partial", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("wm_tiled", C.c_void_p),
        ("adj_grad_phase", C.c_void_p),
        ("adj_cot_scale", C.c_float),
        ("loss_target_u8", C.c_int32),
    ]


class AsmError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"{STATUS_NAMES.get(code, code)}: {message}")
        self.code = code


_lib = None
_lock = threading.Lock()


def load():
    """Load libasm_b200.so (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  learned_hologram_gan_b200 has no CPU or PyTorch fallback."
            )
        lib = C.CDLL(LIB_PATH)
        lib.asm_version.restype = C.c_int
        lib.asm_sizeof_io.restype = C.c_int
        lib.asm_last_error.restype = C.c_char_p
        lib.asm_plan_create.restype = C.c_int
        lib.asm_plan_create.argtypes = [
            C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_double, C.POINTER(C.c_float), C.c_int, C.c_double,
        ]
        lib.asm_plan_destroy.restype = C.c_int
        lib.asm_plan_destroy.argtypes = [C.c_void_p]
        lib.asm_plan_info.restype = C.c_int
        lib.asm_plan_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int]
        lib.asm_workspace_bytes.restype = C.c_size_t
        lib.asm_workspace_bytes.argtypes = [C.c_void_p, C.POINTER(AsmIO)]
        lib.asm_build_grid.restype = C.c_int
        lib.asm_build_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.asm_wm_tiled_bytes.restype = C.c_size_t
        lib.asm_wm_tiled_bytes.argtypes = [C.c_void_p]
        lib.asm_build_wm_tiled.restype = C.c_int
        lib.asm_build_wm_tiled.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.asm_propagate.restype = C.c_int
        lib.asm_propagate.argtypes = [C.c_void_p, C.POINTER(AsmIO), C.c_void_p]
        lib.asm_launch_count.restype = C.c_longlong
        lib.asm_loss_finish.restype = C.c_int
        lib.asm_loss_finish.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        lib.asm_fused_step_supported.restype = C.c_int
        lib.asm_fused_step_supported.argtypes = [C.c_void_p]
        lib.asm_profile_enable.restype = C.c_int
        lib.asm_profile_enable.argtypes = [C.c_int]
        lib.asm_profile_collect.restype = C.c_int
        lib.asm_profile_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]
        if lib.asm_sizeof_io() != C.sizeof(AsmIO):
            raise ImportError(
                f"asm_io layout mismatch: library {lib.asm_sizeof_io()} bytes, binding {C.sizeof(AsmIO)} bytes")
        _lib = lib
    return _lib


def check(code):
    if code != 0:
        raise AsmError(code, load().asm_last_error().decode("utf-8", "replace"))
