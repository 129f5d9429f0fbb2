"""The C-ABI library loads without a GPU and exports exactly what include/pgt_scan.h and include/pgt_extreme.h declare;
compute entry points fail loudly (PGT_ERR_CUDA) when no device is usable -- no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for hdr in ("pgt_scan.h", "pgt_extreme.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(pgt_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_every_declared_symbol_is_exported_and_bound():
    from popgenomicstools_b200 import _cabi
    lib = _cabi.load()
    decl = declared_functions()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/*.h but not exported by libpgtscan.so"
        assert name in _cabi.PROTOTYPES, f"{name} has no ctypes prototype in _cabi.py"
    for name in _cabi.PROTOTYPES:
        assert name in decl, f"{name} bound in _cabi.py but not declared in include/*.h"
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
    xhdr = open(os.path.join(ROOT, "include", "pgt_extreme.h")).read()
    xw = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\} pgt_xwindows;", xhdr, re.S).group(1)
    xw = re.sub(r"/\*.*?\*/", "", xw, flags=re.S)
    assert tuple(re.findall(r"\*\s*(\w+);", xw)) == _cabi.XWINDOW_FIELDS
    assert C.sizeof(_cabi.PgtXWindows) == 8 * len(_cabi.XWINDOW_FIELDS) == 48


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
    # extreme scan: the host bookkeeping works without a device, the scan itself does not
    pos = np.arange(1, 101, dtype=np.uint32)
    xh = C.c_void_p()
    assert lib.pgt_xplan_create(C.byref(xh), pos.ctypes.data, off.ctypes.data, None, 1, 10, 0) == 0
    assert lib.pgt_xplan_num_windows(xh) == 11
    xo = _cabi.PgtXWindows()
    val = np.zeros(11)
    xo.ext_value = val.ctypes.data
    rc = lib.pgt_scan_extreme(xh, None, 0, 2.0, pos.ctypes.data, a.ctypes.data, C.byref(xo), ws.ctypes.data, ws.nbytes,
                              _cabi.PGT_MEM_HOST, None)
    assert rc == _cabi.PGT_ERR_CUDA
    # the round-2 entry points: one-process multi-GPU scan, streaming uploader, IPC export -- none has a CPU path
    dev = (C.c_int * 2)(0, 1)
    xo2 = _cabi.PgtXWindows()
    xo2.ext_value = val.ctypes.data
    assert lib.pgt_scan_extreme_sharded(xh, 0, 2.0, pos.ctypes.data, a.ctypes.data, C.byref(xo2), dev, 2) == _cabi.PGT_ERR_CUDA
    lib.pgt_xplan_destroy(xh)
    h2 = C.c_void_p()
    assert lib.pgt_plan_create(C.byref(h2), 0, off.ctypes.data, 1, 10, 5, 0) == 0
    so, sb = np.zeros(19), np.zeros(19)
    out2 = _cabi.PgtWindows()
    out2.sum_a, out2.sum_b = so.ctypes.data, sb.ctypes.data
    rc = lib.pgt_scan_sharded(h2, _cabi.PGT_STAT_FST, C.byref(cols), 1, None, C.byref(out2), dev, 2, None, None)
    assert rc == _cabi.PGT_ERR_CUDA and b"no CPU fallback" in lib.pgt_last_error()
    assert lib.pgt_scan_sharded(h2, _cabi.PGT_STAT_FST, C.byref(cols), 1, None, C.byref(out2), None, 0, None, None) == _cabi.PGT_ERR_ARGS
    assert lib.pgt_plan_scan_path(h2, _cabi.PGT_STAT_FST) == 0 and lib.pgt_plan_scan_path(h2, 9) == _cabi.PGT_ERR_ARGS
    lib.pgt_plan_destroy(h2)
    up = C.c_void_p()
    assert lib.pgt_uploader_create(C.byref(up), None, 0, 4, 1 << 20, 2) == _cabi.PGT_ERR_CUDA and not up.value
    assert lib.pgt_uploader_create(C.byref(up), None, 0, 0, 1 << 20, 2) == _cabi.PGT_ERR_ARGS
    handle = (C.c_ubyte * 64)()
    assert lib.pgt_ipc_export(ws.ctypes.data, handle, 64) == _cabi.PGT_ERR_CUDA
    assert lib.pgt_ipc_export(ws.ctypes.data, handle, 8) == _cabi.PGT_ERR_ARGS
    fb, tb = C.c_size_t(), C.c_size_t()
    assert lib.pgt_device_mem_info(C.byref(fb), C.byref(tb)) == _cabi.PGT_ERR_CUDA


def test_tune_knobs_and_the_env_hook():
    """pgt_tune accepts the documented knobs and rejects others; PGT_TUNE applies knobs when the
    Python layer loads the library (forced-path test runs: `PGT_TUNE=slide=2 pytest -m gpu`)."""
    import subprocess
    import sys
    from popgenomicstools_b200 import _cabi
    lib = _cabi.load()
    for key, good, bad in ((b"level1", 2, None), (b"level2", 2, 3), (b"slide", 2, 3), (b"stages", 3, 9),
                           (b"stage_kb", 64, 500), (b"xgroup", 8, 5), (b"xsmall", 2, 3), (b"persite", 1, 2), (b"fused2", 1, 2), (b"slideglobal", 1, 2),
                           (b"unittable", 2, 3)):
        assert lib.pgt_tune(key, good) == 0, key
        if bad is not None:
            assert lib.pgt_tune(key, bad) == _cabi.PGT_ERR_ARGS, key
    assert lib.pgt_tune(b"nonsense", 1) == _cabi.PGT_ERR_ARGS
    for key, dflt in ((b"level1", 0), (b"level2", 0), (b"slide", 0), (b"stages", 2), (b"stage_kb", 110), (b"xgroup", 0), (b"xsmall", 0),
                      (b"persite", 0), (b"fused2", 0), (b"slideglobal", 0), (b"unittable", 0)):
        assert lib.pgt_tune(key, dflt) == 0
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    run = lambda v: subprocess.run([sys.executable, "-c", "import popgenomicstools_b200"], cwd=root,
                                   env=dict(os.environ, PGT_TUNE=v), capture_output=True, text=True)
    assert run("slide=1,level2=1").returncode == 0
    r = run("bogus=1")
    assert r.returncode != 0 and "PGT_TUNE" in r.stderr
