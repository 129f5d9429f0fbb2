"""The C-ABI library loads without a GPU and exports exactly what include/pgt_scan.h declares;
compute entry points fail loudly (PGT_ERR_CUDA) when no device is usable -- no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "pgt_scan.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(pgt_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported_and_bound():
    from popgenomicstools_b200 import _cabi
    lib = _cabi.load()
    decl = declared_functions()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in pgt_scan.h but not exported by libpgtscan.so"
        assert name in _cabi.PROTOTYPES, f"{name} has no ctypes prototype in _cabi.py"
    for name in _cabi.PROTOTYPES:
        assert name in decl, f"{name} bound in _cabi.py but not declared in pgt_scan.h"
    assert lib.pgt_abi_version() == 1


def test_struct_layouts_match_header():
    from popgenomicstools_b200 import _cabi
    assert C.sizeof(_cabi.PgtRange) == 32
    assert C.sizeof(_cabi.PgtColumns) == 8 * len(_cabi.COLUMN_FIELDS) == 64
    assert C.sizeof(_cabi.PgtWindows) == 8 * len(_cabi.WINDOW_FIELDS) == 120
    hdr = open(os.path.join(ROOT, "include", "pgt_scan.h")).read()
    cols = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\} pgt_columns;", hdr, re.S).group(1)
    cols = re.sub(r"/\*.*?\*/", "", cols, flags=re.S)
    assert tuple(re.findall(r"\*\s*(\w+);", cols)) == _cabi.COLUMN_FIELDS
    wins = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\} pgt_windows;", hdr, re.S).group(1)
    wins = re.sub(r"/\*.*?\*/", "", wins, flags=re.S)
    assert tuple(re.findall(r"\*\s*(\w+);", wins)) == _cabi.WINDOW_FIELDS


def test_scan_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from popgenomicstools_b200 import _cabi
    lib = _cabi.load()
    off = np.array([0, 100], np.uint64)
    h = C.c_void_p()
    assert lib.pgt_plan_create(C.byref(h), 0, off.ctypes.data, 1, 10, 5, 0) == 0
    a = np.zeros(100)
    cols = _cabi.PgtColumns()
    cols.a = a.ctypes.data
    cols.b = a.ctypes.data
    out = _cabi.PgtWindows()
    ws = np.zeros(1 << 20, np.uint8)
    rc = lib.pgt_scan_fst(h, None, C.byref(cols), C.byref(out), ws.ctypes.data, ws.nbytes, _cabi.PGT_MEM_HOST, None)
    assert rc == _cabi.PGT_ERR_CUDA
    assert b"CUDA" in lib.pgt_last_error()
    assert lib.pgt_synth_fst(1, 0, 10, a.ctypes.data, a.ctypes.data, None) < 0
    lib.pgt_plan_destroy(h)
