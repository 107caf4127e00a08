"""The C-ABI library loads without a GPU and exports every symbol include/asm_b200.h declares."""

import ctypes
import os
import re

from conftest import ROOT

from learned_hologram_gan_b200 import _cabi


def declared_functions():
    text = open(os.path.join(ROOT, "include", "asm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(asm_[a-z_]+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_functions()
    assert names, "no declarations parsed"
    assert sorted(_cabi.EXPORTS) == names


def test_library_loads_and_exports_every_symbol():
    lib = _cabi.load()
    for name in declared_functions():
        assert hasattr(lib, name), name
    header = open(os.path.join(ROOT, "include", "asm_b200.h")).read()
    version = int(re.search(r"#define ASM_B200_VERSION (\d+)", header).group(1))
    assert lib.asm_version() == version
    assert lib.asm_sizeof_io() == ctypes.sizeof(_cabi.AsmIO)
    assert isinstance(lib.asm_last_error(), bytes)
    assert lib.asm_launch_count() == 0


def test_descriptor_is_validated_without_touching_a_device():
    lib = _cabi.load()
    io = _cabi.AsmIO()
    io.struct_bytes = 12  # wrong on purpose
    assert lib.asm_workspace_bytes(None, ctypes.byref(io)) == 0
    assert b"null plan" in lib.asm_last_error()
    assert lib.asm_propagate(None, ctypes.byref(io), None) == -1  # ASM_EINVAL, no launch
    assert lib.asm_launch_count() == 0


def test_enums_match_header():
    text = open(os.path.join(ROOT, "include", "asm_b200.h")).read()
    for name, value in re.findall(r"\b(ASM_[A-Z0-9_]+)\s*=\s*(-?\d+)", text):
        short = name[len("ASM_"):]
        if hasattr(_cabi, short):
            assert getattr(_cabi, short) == int(value), name
        elif short.startswith("FILTER_") and hasattr(_cabi, "FLAG_" + short[len("FILTER_"):]):
            assert getattr(_cabi, "FLAG_" + short[len("FILTER_"):]) == int(value), name
