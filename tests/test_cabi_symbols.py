"""The C-ABI library loads and exports exactly what include/mcz_b200.h declares (no GPU needed)."""
import os
import re
import subprocess

import pytest

from pymcz_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions(name="mcz_b200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcz_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _header_functions() == sorted(_lib.SIGNATURES)
    assert _header_functions("mcz_b200_debug.h") == sorted(_lib.DEBUG_SIGNATURES)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (mcz_[a-z0-9_]+)", out))
    # every exported mcz_* symbol is declared: the boundary in mcz_b200.h, test hooks in mcz_b200_debug.h
    assert set(_header_functions()) | set(_header_functions("mcz_b200_debug.h")) == exported
    L = _lib.lib()       # loads, binds every symbol
    assert L.mcz_version().startswith(b"pymcz_b200")
    from pymcz_b200._scales import KEYS, LINES
    assert [L.mcz_key_name(i).decode() for i in range(len(KEYS))] == KEYS
    assert [L.mcz_line_name(i).decode() for i in range(len(LINES))] == LINES


def test_sm100a_code_is_present():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_oracle_import_in_product():
    """The product path must never route through the oracle (or any CPU fallback)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pymcz_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "mcz_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_ops_refuse_cpu_tensors():
    import torch
    from pymcz_b200 import ops
    with pytest.raises(ValueError):
        ops.forward_production(torch.zeros(2, 11), torch.zeros(2, 11), 100, 1)
