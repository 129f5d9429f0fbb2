"""Host-side mirror of the reference's window tools over the C ABI (include/pgt_scan.h).

The reference exposes the path only as CLIs; argument names follow them:
``winsize`` / ``stepsize`` as in /root/reference/fstWindow.cpp:158-162 (number of SITES) and
dxyWindow's ``-winsize -stepsize -minind -fixedsite -skip_missing`` (dxyWindow.cpp:34-61).

PyTorch is plumbing only: device memory for columns/outputs/workspace and the current CUDA
stream.  All arithmetic happens in libpgtscan.so's kernels; nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import PgtFstOut, PgtRange, check


def _torch():
    import torch
    return torch


class WindowPlan:
    """Closed-form window enumeration (pgt_plan_*): which windows the reference would print.

    contig_offsets: cumulative site counts (mode='sites') or chromosome lengths in bp
    (mode='bp'), length ncontig+1, in file order.
    """

    def __init__(self, contig_offsets, winsize=1, stepsize=1, mode="sites", unit_sites=0):
        lib = _cabi.load()
        self._lib = lib
        self.offsets = np.ascontiguousarray(contig_offsets, dtype=np.uint64)
        if self.offsets.ndim != 1 or len(self.offsets) < 1:
            raise ValueError("contig_offsets must be a 1-D array of length ncontig+1")
        self.winsize, self.stepsize = int(winsize), int(stepsize)
        self.mode = {"sites": _cabi.PGT_MODE_SITES, "bp": _cabi.PGT_MODE_BP}[mode]
        if not (0 <= self.winsize < 2**32 and 0 <= self.stepsize < 2**32):
            raise _cabi.PgtError(_cabi.PGT_ERR_ARGS, "winsize/stepsize out of range")
        h = C.c_void_p()
        check(lib.pgt_plan_create(C.byref(h), self.mode, self.offsets.ctypes.data, len(self.offsets) - 1,
                                  self.winsize, self.stepsize, int(unit_sites)))
        self._h = h
        self._workspaces = {}

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.pgt_plan_destroy(h)

    @property
    def handle(self):
        return self._h

    @property
    def num_windows(self):
        return int(self._lib.pgt_plan_num_windows(self._h))

    @property
    def num_units(self):
        return int(self._lib.pgt_plan_num_units(self._h))

    @property
    def num_segments(self):
        return int(self._lib.pgt_plan_num_segments(self._h))

    @property
    def num_sites(self):
        return int(self._lib.pgt_plan_num_sites(self._h))

    def windows(self):
        """(first, last, label) arrays: inclusive site/entry indices and contig index per window."""
        n = self.num_windows
        first, last, label = np.empty(n, np.uint64), np.empty(n, np.uint64), np.empty(n, np.uint32)
        check(self._lib.pgt_plan_windows(self._h, first.ctypes.data, last.ctypes.data, label.ctypes.data))
        return first, last, label

    def unit(self, j):
        st, ln = C.c_uint64(), C.c_uint32()
        check(self._lib.pgt_plan_unit(self._h, j, C.byref(st), C.byref(ln)))
        return st.value, ln.value

    def window_units(self, w):
        f, c = C.c_uint64(), C.c_uint64()
        check(self._lib.pgt_plan_window_units(self._h, w, C.byref(f), C.byref(c)))
        return f.value, c.value

    def shard(self, rank, nranks):
        """Window range [w_lo, w_hi) and site range [site_lo, site_hi) of shard `rank`."""
        v = [C.c_uint64() for _ in range(4)]
        check(self._lib.pgt_plan_shard(self._h, rank, nranks, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def workspace_bytes(self, window_range=None, site_origin=0):
        r = _range(self, window_range, site_origin)
        return int(self._lib.pgt_scan_workspace_bytes(self._h, C.byref(r)))

    def workspace(self, device, window_range=None, site_origin=0):
        """Device scratch for a scan, cached per (device, range)."""
        torch = _torch()
        key = (str(device), window_range, site_origin)
        ws = self._workspaces.get(key)
        if ws is None:
            ws = torch.empty(self.workspace_bytes(window_range, site_origin), dtype=torch.uint8, device=device)
            self._workspaces[key] = ws
        return ws


def _range(plan, window_range, site_origin):
    lo, hi = (0, plan.num_windows) if window_range is None else window_range
    return PgtRange(int(lo), int(hi), int(site_origin))


def _dev_ptr(t, dtype, name):
    torch = _torch()
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor")
    if t.dtype != dtype or not t.is_contiguous():
        raise TypeError(f"{name}: expected contiguous {dtype}, got {t.dtype}")
    return t.data_ptr()


def fst_window(plan, pos, a, b, window_range=None, site_origin=0, out=None):
    """Sliding-window FST = sum(a)/sum(b) (fstWindow.cpp:69-107) over device columns.

    pos: uint32/int32 CUDA tensor (may be None: positions are then not gathered), a, b: float64
    CUDA tensors holding global sites [site_origin, ...).  Returns a dict of CUDA tensors, one
    element per window of `window_range` (default: all windows): label, start_pos, end_pos,
    mid_pos, sum_a, sum_b, fst, nsites.  Enqueued on the current stream, not synchronised.
    """
    torch = _torch()
    lib = plan._lib
    dev = a.device
    r = _range(plan, window_range, site_origin)
    nwin = r.w_hi - r.w_lo
    if out is None:
        out = {k: torch.empty(nwin, dtype=torch.uint32, device=dev) for k in ("label", "start_pos", "end_pos", "mid_pos", "nsites")}
        out.update({k: torch.empty(nwin, dtype=torch.float64, device=dev) for k in ("sum_a", "sum_b", "fst")})
    if pos is not None and pos.dtype == torch.int32:
        pos = pos.view(torch.uint32)
    o = PgtFstOut(*[out[k].data_ptr() if k in out and out[k] is not None else None for k, _ in PgtFstOut._fields_])
    ws = plan.workspace(dev, window_range, site_origin)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.pgt_scan_fst(plan.handle, C.byref(r), _dev_ptr(pos, torch.uint32, "pos"), _dev_ptr(a, torch.float64, "a"),
                               _dev_ptr(b, torch.float64, "b"), C.byref(o), ws.data_ptr(), ws.numel(), _cabi.PGT_MEM_DEVICE,
                               C.c_void_p(stream)))
    return out


# ---- synthetic inputs (device) -------------------------------------------------------------

def synth_fst(seed, site0, n, device="cuda"):
    torch = _torch()
    lib = _cabi.load()
    a = torch.empty(n, dtype=torch.float64, device=device)
    b = torch.empty(n, dtype=torch.float64, device=device)
    with torch.cuda.device(a.device):
        st = torch.cuda.current_stream(a.device).cuda_stream
        check(lib.pgt_synth_fst(seed, site0, n, a.data_ptr(), b.data_ptr(), C.c_void_p(st)))
    return a, b


def synth_pos(seed, site0, n, contig_offsets, density=1, device="cuda"):
    torch = _torch()
    lib = _cabi.load()
    off = np.ascontiguousarray(contig_offsets, dtype=np.uint64)
    pos = torch.empty(n, dtype=torch.uint32, device=device)
    with torch.cuda.device(pos.device):
        st = torch.cuda.current_stream(pos.device).cuda_stream
        check(lib.pgt_synth_pos(seed, site0, n, off.ctypes.data, len(off) - 1, density, pos.data_ptr(), C.c_void_p(st)))
    return pos
