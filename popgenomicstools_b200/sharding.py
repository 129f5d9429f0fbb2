"""Multi-GPU plumbing: one process per GPU, shards of the window list, gather to rank 0.

The scan itself needs no collective: shard r holds sites
[site_lo, site_hi) -- its windows plus a halo of W-S sites shared with the next shard -- and
computes windows [w_lo, w_hi) independently.  Only the per-window results travel: every rank
keeps its result columns in ONE packed buffer (SoA, padded to the largest shard) so the gather
to rank 0 is a single torch.distributed call (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
import numpy as np

F64_FIELDS = ("sum_a", "sum_b", "fst", "het", "dxy", "ext_value", "prop")  # every other field is uint32


class PackedWindows:
    """Result columns of one shard as views into a single uint8 buffer."""

    def __init__(self, fields, nwin_local, maxwin, device, torch):
        self.torch = torch
        self.fields = [k for k in fields if k != "dxy_global"]
        self.has_global = "dxy_global" in fields
        self.nwin_local, self.maxwin = int(nwin_local), max(int(maxwin), 1)
        self.f64 = [k for k in self.fields if k in F64_FIELDS]
        self.u32 = [k for k in self.fields if k not in F64_FIELDS]
        nbytes = self.maxwin * (8 * len(self.f64) + 4 * len(self.u32)) + 24 + 40
        nbytes = (nbytes + 7) // 8 * 8
        self.buffer = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.views = self._views(self.buffer, self.nwin_local)

    def _views(self, buf, n):
        torch = self.torch
        out, o = {}, 0
        for k in self.f64:
            out[k] = buf[o:o + 8 * self.maxwin].view(torch.float64)[:n]
            o += 8 * self.maxwin
        for k in self.u32:
            out[k] = buf[o:o + 4 * self.maxwin].view(torch.uint32)[:n]
            o += 4 * self.maxwin
        o = (o + 7) // 8 * 8
        if self.has_global:
            out["dxy_global"] = buf[o:o + 24].view(torch.float64)
        return out

    def gather(self, dist, rank, world, gather_list=None):
        """Gather every rank's packed buffer to rank 0 (one collective).  Returns the list on
        rank 0 (reuse it across steps via `gather_list`), None elsewhere."""
        if world == 1:
            return [self.buffer]
        if rank == 0 and gather_list is None:
            gather_list = [self.torch.empty_like(self.buffer) for _ in range(world)]
        dist.gather(self.buffer, gather_list if rank == 0 else None, dst=0)
        return gather_list if rank == 0 else None

    def unpack(self, gathered, counts):
        """Rank 0: concatenate the shards' columns in window order -> dict of numpy arrays."""
        res = {k: [] for k in self.fields}
        glob = np.zeros(3)
        for buf, n in zip(gathered, counts):
            v = self._views(buf, int(n))
            for k in self.fields:
                res[k].append(v[k].cpu().numpy())
            if self.has_global:
                g = v["dxy_global"].cpu().numpy()
                glob += g  # shards own disjoint unit ranges of the global line
        out = {k: np.concatenate(v) for k, v in res.items()}
        if self.has_global:
            out["dxy_global"] = glob
        return out


def shard_counts(dist, nwin_local, world, device, torch):
    """All ranks' window counts (needed to size the packed buffers)."""
    if world == 1:
        return [int(nwin_local)]
    t = torch.tensor([int(nwin_local)], device=device, dtype=torch.int64)
    allc = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allc, t)
    return [int(x.item()) for x in allc]


# ---- one table in rank 0's memory, written in place by every rank -----------------------------------------------

_NP = {"f8": np.float64, "u4": np.uint32}


class RawDeviceArray:
    """A device array given by pointer (memory from pgt_device_alloc, possibly mapped from another process).
    scan(..., out={...}) accepts it wherever it accepts a CUDA tensor; torch.as_tensor() views it on the owner."""

    def __init__(self, ptr, n, kind):
        self.ptr, self.n, self.kind = int(ptr), int(n), kind
        self.__cuda_array_interface__ = {"shape": (self.n,), "typestr": "<" + kind, "data": (self.ptr, False), "version": 2}

    def data_ptr(self):
        return self.ptr


class SharedTable:
    """Where the window table of a multi-rank scan lives.  Small tables (the headline C4 table is 13 MB) live ONCE,
    on rank 0: SoA, one array of `nwin_total` elements per field, plus one dxy_global triple per rank; rank r
    receives pointers to rows [w_lo, w_hi) of every array and hands them to the scan as its output arrays, so its
    window kernel stores the rows where they finally belong -- over NVLink for r > 0 (CUDA IPC mapping of rank 0's
    allocation, pgt_ipc_*).  Nothing is gathered; after a barrier rank 0 reads the table.  Tables above
    `rank0_limit` bytes (one row per site: GBs) stay sharded, every rank keeping its own rows in its own HBM:
    pushing them through one GPU's NVLink port every step would time the port, not the scan.  Either way
    checksum() is the same function of the table's content, whatever the number of ranks.  backend "shm" does the
    rank-0 placement with a file in /dev/shm and numpy views: the CPU (gloo) tests of this layout logic."""

    def __init__(self, fields, nwin_total, rank, world, dist, torch=None, device=None, backend="cuda-ipc", tag="pgt",
                 w_lo=0, w_hi=None, rank0_limit=256 << 20):
        self.fields = [k for k in fields if k != "dxy_global"]
        self.has_global = "dxy_global" in fields
        self.nwin, self.rank, self.world, self.dist, self.torch, self.device = int(nwin_total), rank, world, dist, torch, device
        self.backend = backend
        self.w_lo, self.w_hi = int(w_lo), int(self.nwin if w_hi is None else w_hi)
        self.kind = {k: ("f8" if k in F64_FIELDS else "u4") for k in self.fields}
        row_bytes = sum(8 if self.kind[k] == "f8" else 4 for k in self.fields)
        self.placement = "rank0" if (world == 1 or backend != "cuda-ipc" or row_bytes * self.nwin <= rank0_limit) else "sharded"
        self.row0 = 0 if self.placement == "rank0" else self.w_lo          # global window index of row 0 of this process' arrays
        nrows = self.nwin if self.placement == "rank0" else self.w_hi - self.w_lo
        self.off, o = {}, 0
        for k in self.fields:
            self.off[k] = o
            o = (o + (8 if self.kind[k] == "f8" else 4) * max(nrows, 1) + 255) // 256 * 256
        self.goff = o
        self.nbytes = o + 24 * world + 256
        self.base = None
        self._mapped = False
        if backend == "cuda-ipc":
            self._open_cuda()
        else:
            self._open_shm(tag)

    # -- CUDA: rank 0 allocates and exports, the others map (rank-0 placement); or everyone allocates its own rows
    def _open_cuda(self):
        import ctypes as C
        from . import _cabi
        torch, lib = self.torch, _cabi.load()
        owner = self.rank == 0 or self.placement == "sharded"
        if owner:
            p = C.c_void_p()
            _cabi.check(lib.pgt_device_alloc(C.byref(p), self.nbytes))
            self.base = p.value
        if self.world > 1 and self.placement == "rank0":
            handle = torch.zeros(64, dtype=torch.uint8, device=self.device)
            if self.rank == 0:
                buf = (C.c_ubyte * 64)()
                _cabi.check(lib.pgt_ipc_export(C.c_void_p(self.base), buf, 64))
                handle.copy_(torch.frombuffer(bytearray(buf), dtype=torch.uint8))
            self.dist.broadcast(handle, src=0)
            if self.rank != 0:
                raw = bytes(handle.cpu().numpy().tobytes())
                p = C.c_void_p()
                _cabi.check(lib.pgt_ipc_open(raw, C.byref(p)))
                self.base = p.value
                self._mapped = True

    def _open_shm(self, tag):
        import os
        path = f"/dev/shm/{tag}_{os.environ.get('MASTER_PORT', '0')}.tbl"
        if self.rank == 0:
            with open(path, "wb") as f:
                f.truncate(self.nbytes)
        if self.world > 1:
            self.dist.barrier()
        self._mm = np.memmap(path, dtype=np.uint8, mode="r+", shape=(self.nbytes,))
        self._path = path

    def _array(self, k, lo, hi):
        """rows [lo, hi) (global window indices) of field k"""
        es = 8 if self.kind[k] == "f8" else 4
        lo, hi = lo - self.row0, hi - self.row0
        if self.backend == "cuda-ipc":
            return RawDeviceArray(self.base + self.off[k] + es * lo, hi - lo, self.kind[k])
        return self._mm[self.off[k] + es * lo:self.off[k] + es * hi].view(_NP[self.kind[k]])

    def _global_slot(self):
        slot = self.rank if self.placement == "rank0" else 0
        if self.backend == "cuda-ipc":
            return RawDeviceArray(self.base + self.goff + 24 * slot, 3, "f8")
        return self._mm[self.goff + 24 * slot:self.goff + 24 * slot + 24].view(np.float64)

    def rows(self, w_lo=None, w_hi=None):
        """Output arrays for this rank's rows [w_lo, w_hi) (+ its dxy_global triple)."""
        w_lo = self.w_lo if w_lo is None else int(w_lo)
        w_hi = self.w_hi if w_hi is None else int(w_hi)
        out = {k: self._array(k, w_lo, w_hi) for k in self.fields}
        if self.has_global:
            out["dxy_global"] = self._global_slot()
        return out

    def table(self):
        """Rank 0 of a rank-0 placement, after a barrier: the whole table (CUDA tensors / numpy views)."""
        assert self.rank == 0 and self.placement == "rank0"
        out = {}
        for k in self.fields:
            a = self._array(k, 0, self.nwin)
            out[k] = self.torch.as_tensor(a, device=self.device) if self.backend == "cuda-ipc" else a
        if self.has_global:
            out["dxy_global"] = self.global_line()
        return out

    def global_line(self):
        """dxy_global: the ranks' disjoint partial lines added in rank order (collective when sharded)."""
        if self.backend != "cuda-ipc":
            g = np.array(self._mm[self.goff:self.goff + 24 * self.world].view(np.float64))
            return g.reshape(self.world, 3).sum(axis=0)
        torch = self.torch
        if self.placement == "rank0":
            if self.rank != 0:
                return None
            g = torch.as_tensor(RawDeviceArray(self.base + self.goff, 3 * self.world, "f8"), device=self.device).cpu().numpy()
            return g.reshape(self.world, 3).sum(axis=0)
        mine = torch.as_tensor(self._global_slot(), device=self.device).clone()
        parts = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(parts, mine)
        return np.sum([p.cpu().numpy() for p in parts], axis=0)

    def checksum(self):
        """Collective (call after a barrier).  Order-sensitive 64-bit checksum per field, sum over all windows i of
        word_i * (2 i + 1) mod 2^64 with i the GLOBAL window index, as one hex string: every rank sums the rows it
        wrote, the partial sums are added.  Equal tables -- bit for bit -- give equal strings, whatever the number
        of ranks that wrote them and wherever the rows live."""
        lo, hi = self.w_lo, self.w_hi
        accs = []
        for k in self.fields:
            a = self._array(k, lo, hi)
            if self.backend == "cuda-ipc":
                torch = self.torch
                t = torch.as_tensor(a, device=self.device) if hi > lo else torch.zeros(0, dtype=torch.float64, device=self.device)
                w = t.view(torch.int64) if self.kind[k] == "f8" else t.view(torch.int32).to(torch.int64)
                acc, step = 0, 1 << 26
                for i in range(0, w.numel(), step):  # bounded temporaries for tables of 1e8 rows
                    x = w[i:i + step]
                    idx = torch.arange(lo + i, lo + i + x.numel(), device=x.device, dtype=torch.int64) * 2 + 1
                    acc = (acc + int((x * idx).sum().item())) & 0xFFFFFFFFFFFFFFFF
            else:
                w = a.view(np.int64) if self.kind[k] == "f8" else a.view(np.int32).astype(np.int64)
                with np.errstate(over="ignore"):
                    acc = int((w * (np.arange(lo, hi, dtype=np.int64) * 2 + 1)).sum()) & 0xFFFFFFFFFFFFFFFF
            accs.append(acc)
        if self.world > 1:
            if self.backend == "cuda-ipc":
                t = self.torch.tensor([a - (1 << 64) if a >= (1 << 63) else a for a in accs], dtype=self.torch.int64, device=self.device)
            else:
                t = self.torch.tensor([a - (1 << 64) if a >= (1 << 63) else a for a in accs], dtype=self.torch.int64)
            self.dist.all_reduce(t)  # int64 sums wrap: addition mod 2^64
            accs = [int(x) & 0xFFFFFFFFFFFFFFFF for x in t.tolist()]
        return "-".join(f"{a:016x}" for a in accs)

    def close(self):
        """Collective: unmap on the other ranks, then the owners free."""
        import ctypes as C
        if self.backend == "cuda-ipc":
            from . import _cabi
            lib = _cabi.load()
            if self._mapped:
                _cabi.check(lib.pgt_ipc_close(C.c_void_p(self.base)))
            if self.world > 1:
                self.dist.barrier()
            if not self._mapped and self.base:
                _cabi.check(lib.pgt_device_free(C.c_void_p(self.base)))
            self.base = None
        else:
            import os
            del self._mm
            if self.world > 1:
                self.dist.barrier()
            if self.rank == 0:
                os.remove(self._path)


def table_checksum(table, fields):
    """SharedTable.checksum() of a whole table given as numpy arrays (the end-to-end path's host table)."""
    parts = []
    for k in fields:
        if k == "dxy_global":
            continue
        a = np.ascontiguousarray(table[k])
        w = a.view(np.int64) if a.dtype == np.float64 else a.view(np.int32).astype(np.int64)
        with np.errstate(over="ignore"):
            acc = int((w * (np.arange(w.size, dtype=np.int64) * 2 + 1)).sum()) & 0xFFFFFFFFFFFFFFFF
        parts.append(f"{acc:016x}")
    return "-".join(parts)
