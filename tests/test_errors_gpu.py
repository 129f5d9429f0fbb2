"""Error behaviour of the C ABI on a live device: bad arguments are reported, never crash."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_argument_and_workspace_errors():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import popgenomicstools_b200 as pgt
    from popgenomicstools_b200 import _cabi
    lib = _cabi.load()
    plan = pgt.WindowPlan([0, 1000], 100, 10)
    a = torch.zeros(1000, dtype=torch.float64, device="cuda")
    cols = _cabi.PgtColumns()
    cols.a = a.data_ptr()
    out = _cabi.PgtWindows()
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    # missing column b
    rc = lib.pgt_scan_fst(plan.handle, None, C.byref(cols), C.byref(out), ws.data_ptr(), ws.numel(), 0, None)
    assert rc == _cabi.PGT_ERR_ARGS and b"NULL" in lib.pgt_last_error()
    cols.b = a.data_ptr()
    # workspace too small / NULL
    rc = lib.pgt_scan_fst(plan.handle, None, C.byref(cols), C.byref(out), ws.data_ptr(), 16, 0, None)
    assert rc == _cabi.PGT_ERR_NOMEM
    rc = lib.pgt_scan_fst(plan.handle, None, C.byref(cols), C.byref(out), None, 1 << 20, 0, None)
    assert rc == _cabi.PGT_ERR_ARGS
    # window range out of bounds
    r = _cabi.PgtRange(0, plan.num_windows + 1, 0, 0)
    rc = lib.pgt_scan_fst(plan.handle, C.byref(r), C.byref(cols), C.byref(out), ws.data_ptr(), ws.numel(), 0, None)
    assert rc == _cabi.PGT_ERR_ARGS
    # dxy needs minind >= 1; bp plans are dxy-only and need site_offsets
    rc = lib.pgt_scan(plan.handle, None, _cabi.PGT_STAT_DXY, C.byref(cols), 0, None, C.byref(out), ws.data_ptr(), ws.numel(), 0, None)
    assert rc == _cabi.PGT_ERR_ARGS and b"minind" in lib.pgt_last_error()
    bp = pgt.WindowPlan([0, 1000], 100, 10, mode="bp")
    rc = lib.pgt_scan_fst(bp.handle, None, C.byref(cols), C.byref(out), ws.data_ptr(), ws.numel(), 0, None)
    assert rc == _cabi.PGT_ERR_ARGS
    # a valid call still works afterwards, all outputs optional
    rc = lib.pgt_scan_fst(plan.handle, None, C.byref(cols), C.byref(out), ws.data_ptr(), ws.numel(), 0, None)
    assert rc == 0
    torch.cuda.synchronize()
    # empty input: zero windows, nothing launched, no error
    empty = pgt.WindowPlan([0], 5, 1)
    res = pgt.fst_window(empty, None, a[:0], a[:0])
    assert len(res["fst"]) == 0
