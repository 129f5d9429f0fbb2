"""Host-side mirror of ihsWindow / xpehhWindow over the C ABI (include/pgt_extreme.h).

Argument names follow the tools: ``winsize`` (bp, /root/reference/ihsWindow.cpp:47-52),
``cutoff`` (ihsWindow.cpp:54-59; positional for xpehhWindow, xpehhWindow.cpp:55) and ``chrlen``
(the -chrlen table, ihsWindow.cpp:60-65).  PyTorch is plumbing only (device memory, current
stream); the reduction runs in libpgtscan.so's kernels and nothing here computes on the CPU
except the window bookkeeping the library itself does on the host (pgt_xplan_create).
"""
import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import PgtRange, PgtXWindows, check

XWINDOW_FIELDS = _cabi.XWINDOW_FIELDS
_X_DTYPES = {"ext_value": np.float64, "ext_pos": np.uint32, "ext_site": np.uint64, "nbig": np.uint32,
             "nsites": np.uint32, "prop": np.float64}


def _torch():
    import torch
    return torch


class ExtremePlan:
    """Window bookkeeping of ihsWindow / xpehhWindow (pgt_xplan_*): which rows the reference prints
    and which sites each window holds.

    pos: HOST position column (numpy uint32, file order); contig_offsets: cumulative site counts of
    the runs of equal chromosome name; chrlen: per-run -chrlen length (0 = not listed) or None.
    """

    def __init__(self, pos, contig_offsets, winsize=100000, chrlen=None, unit_sites=0):
        lib = _cabi.load()
        self._lib = lib
        self._h = None
        self.pos = np.ascontiguousarray(pos, dtype=np.uint32)
        self.offsets = np.ascontiguousarray(contig_offsets, dtype=np.uint64)
        ncontig = len(self.offsets) - 1
        if self.offsets.ndim != 1 or ncontig < 1:
            raise ValueError("contig_offsets must be a 1-D array of length ncontig+1")
        if len(self.pos) != int(self.offsets[-1]):
            raise ValueError("pos must hold contig_offsets[-1] sites")
        self.chrlen = None if chrlen is None else np.ascontiguousarray(chrlen, dtype=np.uint32)
        if self.chrlen is not None and len(self.chrlen) != ncontig:
            raise ValueError("chrlen must have one entry per contig run")
        if not 0 <= int(winsize) < 2**32:
            raise _cabi.PgtError(_cabi.PGT_ERR_ARGS, "winsize out of range")
        self.winsize = int(winsize)
        h = C.c_void_p()
        check(lib.pgt_xplan_create(C.byref(h), self.pos.ctypes.data, self.offsets.ctypes.data,
                                   self.chrlen.ctypes.data if self.chrlen is not None else None, ncontig, self.winsize,
                                   int(unit_sites)))
        self._h = h
        self._workspaces = {}
        self._bound = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.pgt_xplan_destroy(h)

    @property
    def handle(self):
        return self._h

    @property
    def num_windows(self):
        return int(self._lib.pgt_xplan_num_windows(self._h))

    @property
    def num_units(self):
        return int(self._lib.pgt_xplan_num_units(self._h))

    @property
    def num_sites(self):
        return int(self._lib.pgt_xplan_num_sites(self._h))

    def windows(self):
        """dict(label, start, end, first_site, nsites): the printed window table (numpy)."""
        n = self.num_windows
        o = dict(label=np.empty(n, np.uint32), start=np.empty(n, np.uint32), end=np.empty(n, np.uint32),
                 first_site=np.empty(n, np.uint64), nsites=np.empty(n, np.uint32))
        check(self._lib.pgt_xplan_windows(self._h, *[o[k].ctypes.data for k in ("label", "start", "end", "first_site", "nsites")]))
        return o

    def shard(self, rank, nranks):
        """Window range [w_lo, w_hi) and site range [site_lo, site_hi) of shard `rank` (no halo)."""
        v = [C.c_uint64() for _ in range(4)]
        check(self._lib.pgt_xplan_shard(self._h, rank, nranks, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def bind(self, device):
        """Keep the window tables resident on `device` (pgt_xplan_bind_device); later scans on that
        device skip the per-call table upload."""
        torch = _torch()
        device = torch.device(device)
        if self._bound is not None and self._bound[0] == device:
            return
        buf = torch.empty(int(self._lib.pgt_xplan_device_bytes(self._h)), dtype=torch.uint8, device=device)
        with torch.cuda.device(device):
            check(self._lib.pgt_xplan_bind_device(self._h, buf.data_ptr(), buf.numel(),
                                                  C.c_void_p(torch.cuda.current_stream(device).cuda_stream)))
        self._bound = (device, buf)
        self._workspaces = {}

    def workspace(self, device, mem, window_range=None, site_origin=0):
        torch = _torch()
        key = (str(device), mem, window_range, site_origin)
        ws = self._workspaces.get(key)
        if ws is None:
            r = _xrange(self, window_range, site_origin)
            ws = torch.empty(int(self._lib.pgt_scan_extreme_workspace_bytes(self._h, C.byref(r), mem)), dtype=torch.uint8,
                             device=device)
            self._workspaces[key] = ws
        return ws


def _xrange(plan, window_range, site_origin):
    lo, hi = (0, plan.num_windows) if window_range is None else window_range
    return PgtRange(int(lo), int(hi), int(site_origin), 0)


def scan_extreme(plan, stat, cutoff, pos, score, window_range=None, site_origin=0, out=None, device=None):
    """pgt_scan_extreme.  CUDA tensors select PGT_MEM_DEVICE (enqueued on the current stream, CUDA
    tensors back, not synchronised); numpy arrays select PGT_MEM_HOST (numpy arrays back)."""
    torch = _torch()
    lib = plan._lib
    on_device = isinstance(score, torch.Tensor) and score.is_cuda
    r = _xrange(plan, window_range, site_origin)
    nwin = r.w_hi - r.w_lo
    keep = []

    def ptr(v, want, name):
        if v is None:
            return None
        if on_device:
            if not (isinstance(v, torch.Tensor) and v.is_cuda and v.is_contiguous()):
                raise TypeError(f"{name}: expected a contiguous CUDA tensor")
            if name == "pos" and v.dtype == torch.int32:
                v = v.view(torch.uint32)
            if str(v.dtype) != "torch." + np.dtype(want).name:
                raise TypeError(f"{name}: expected {np.dtype(want).name}, got {v.dtype}")
            keep.append(v)
            return v.data_ptr()
        if isinstance(v, torch.Tensor):
            v = v.numpy()
        if v.dtype != np.dtype(want) or not v.flags["C_CONTIGUOUS"]:
            raise TypeError(f"{name}: expected contiguous {np.dtype(want).name}, got {v.dtype}")
        keep.append(v)
        return v.ctypes.data

    p_pos, p_score = ptr(pos, np.uint32, "pos"), ptr(score, np.float64, "score")
    dev = score.device if on_device else torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if out is None:
        out = {}
        for k in XWINDOW_FIELDS:
            if k == "ext_pos" and pos is None:
                continue
            if on_device:
                out[k] = torch.empty(nwin, dtype=getattr(torch, np.dtype(_X_DTYPES[k]).name), device=dev)
            else:
                out[k] = np.empty(nwin, _X_DTYPES[k])
    w = PgtXWindows()
    for k in XWINDOW_FIELDS:
        v = out.get(k)
        if v is not None:
            setattr(w, k, v.data_ptr() if on_device else v.ctypes.data)
    mem = _cabi.PGT_MEM_DEVICE if on_device else _cabi.PGT_MEM_HOST
    if on_device:
        plan.bind(dev)
    elif plan._bound is not None and plan._bound[0] != dev:
        raise ValueError(f"plan is bound to {plan._bound[0]}, scan requested on {dev}")
    ws = plan.workspace(dev, mem, window_range, site_origin)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(lib.pgt_scan_extreme(plan.handle, C.byref(r), int(stat), float(cutoff), p_pos, p_score, C.byref(w), ws.data_ptr(),
                                   ws.numel(), mem, C.c_void_p(stream)))
    return out


def scan_extreme_sharded(plan, stat, cutoff, pos, score, devices):
    """pgt_scan_extreme_sharded: numpy columns over ALL sites, one shard of windows per entry of `devices`
    (CUDA device indices of this process); numpy arrays over all windows back, bit-identical for any list."""
    lib = plan._lib
    pos = None if pos is None else np.ascontiguousarray(pos, np.uint32)
    score = np.ascontiguousarray(score, np.float64)
    out = {k: np.empty(plan.num_windows, _X_DTYPES[k]) for k in XWINDOW_FIELDS if not (k == "ext_pos" and pos is None)}
    w = PgtXWindows()
    for k, v in out.items():
        setattr(w, k, v.ctypes.data)
    dev = (C.c_int * len(devices))(*[int(d) for d in devices])
    check(lib.pgt_scan_extreme_sharded(plan.handle, int(stat), float(cutoff), pos.ctypes.data if pos is not None else None,
                                       score.ctypes.data, C.byref(w), dev, len(devices)))
    return out


def ihs_window(plan, pos, score, cutoff=2.0, **kw):
    """Most extreme |iHS| per bp window and the proportion of |iHS| > cutoff
    (/root/reference/ihsWindow.cpp:93-187).  Returns ext_value (signed score), ext_pos, ext_site,
    nbig, nsites, prop per window; empty windows have nsites 0 and NaN ext_value / prop."""
    return scan_extreme(plan, _cabi.PGT_XSTAT_IHS, cutoff, pos, score, **kw)


def xpehh_window(plan, pos, score, cutoff, **kw):
    """Most extreme XP-EHH per bp window: cutoff < 0 -> minimum and proportion below the cutoff,
    else maximum and proportion above it (/root/reference/xpehhWindow.cpp:87-193)."""
    return scan_extreme(plan, _cabi.PGT_XSTAT_XPEHH, cutoff, pos, score, **kw)


def synth_score(seed, site0, n, device="cuda"):
    torch = _torch()
    s = torch.empty(n, dtype=torch.float64, device=device)
    with torch.cuda.device(s.device):
        check(_cabi.load().pgt_synth_score(seed, site0, n, s.data_ptr(), C.c_void_p(torch.cuda.current_stream(s.device).cuda_stream)))
    return s


def profile_read_extreme():
    """-> dict(units_ms, units_launches, windows_ms, windows_launches) of the extreme scan since the last read."""
    um, wm = C.c_double(), C.c_double()
    un, wn = C.c_uint64(), C.c_uint64()
    check(_cabi.load().pgt_profile_read_extreme(C.byref(um), C.byref(un), C.byref(wm), C.byref(wn)))
    return dict(units_ms=um.value, units_launches=un.value, windows_ms=wm.value, windows_launches=wn.value)
