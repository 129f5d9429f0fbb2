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
    """The WHOLE window table lives once, on rank 0: SoA, one array of `nwin_total` elements per field, plus one
    dxy_global triple per rank.  Rank r receives pointers to rows [w_lo, w_hi) of every array and hands them to the
    scan as its output arrays, so its window kernel stores the rows where they finally belong -- over NVLink for
    r > 0 (CUDA IPC mapping of rank 0's allocation, pgt_ipc_*).  Nothing is gathered; after a barrier rank 0 reads
    the table.  backend "shm" does the same with a file in /dev/shm and numpy views: the CPU (gloo) tests of this
    layout logic."""

    def __init__(self, fields, nwin_total, rank, world, dist, torch=None, device=None, backend="cuda-ipc", tag="pgt"):
        self.fields = [k for k in fields if k != "dxy_global"]
        self.has_global = "dxy_global" in fields
        self.nwin, self.rank, self.world, self.dist, self.torch, self.device = int(nwin_total), rank, world, dist, torch, device
        self.backend = backend
        self.kind = {k: ("f8" if k in F64_FIELDS else "u4") for k in self.fields}
        self.off, o = {}, 0
        for k in self.fields:
            self.off[k] = o
            o = (o + (8 if self.kind[k] == "f8" else 4) * max(self.nwin, 1) + 255) // 256 * 256
        self.goff = o
        self.nbytes = o + 24 * world + 256
        self.base = None
        self._mapped = False
        if backend == "cuda-ipc":
            self._open_cuda()
        else:
            self._open_shm(tag)

    # -- CUDA: rank 0 allocates and exports, the others map
    def _open_cuda(self):
        import ctypes as C
        from . import _cabi
        torch, lib = self.torch, _cabi.load()
        handle = torch.zeros(64, dtype=torch.uint8, device=self.device)
        if self.rank == 0:
            p = C.c_void_p()
            _cabi.check(lib.pgt_device_alloc(C.byref(p), self.nbytes))
            self.base = p.value
            if self.world > 1:
                buf = (C.c_ubyte * 64)()
                _cabi.check(lib.pgt_ipc_export(C.c_void_p(self.base), buf, 64))
                handle.copy_(torch.frombuffer(bytearray(buf), dtype=torch.uint8))
        if self.world > 1:
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
        es = 8 if self.kind[k] == "f8" else 4
        if self.backend == "cuda-ipc":
            return RawDeviceArray(self.base + self.off[k] + es * lo, hi - lo, self.kind[k])
        return self._mm[self.off[k] + es * lo:self.off[k] + es * hi].view(_NP[self.kind[k]])

    def rows(self, w_lo, w_hi):
        """Output arrays for the rows [w_lo, w_hi) (+ this rank's dxy_global triple)."""
        out = {k: self._array(k, int(w_lo), int(w_hi)) for k in self.fields}
        if self.has_global:
            if self.backend == "cuda-ipc":
                out["dxy_global"] = RawDeviceArray(self.base + self.goff + 24 * self.rank, 3, "f8")
            else:
                out["dxy_global"] = self._mm[self.goff + 24 * self.rank:self.goff + 24 * self.rank + 24].view(np.float64)
        return out

    def table(self):
        """Rank 0, after a barrier: the whole table (CUDA tensors / numpy views); dxy_global summed in rank order."""
        assert self.rank == 0
        out = {}
        for k in self.fields:
            a = self._array(k, 0, self.nwin)
            out[k] = self.torch.as_tensor(a, device=self.device) if self.backend == "cuda-ipc" else a
        if self.has_global:
            if self.backend == "cuda-ipc":
                g = self.torch.as_tensor(RawDeviceArray(self.base + self.goff, 3 * self.world, "f8"), device=self.device).cpu().numpy()
            else:
                g = np.array(self._mm[self.goff:self.goff + 24 * self.world].view(np.float64))
            out["dxy_global"] = g.reshape(self.world, 3).sum(axis=0)
        return out

    def checksum(self):
        """Rank 0: order-sensitive 64-bit checksum per field (sum of word_i * (2 i + 1) mod 2^64), one hex string.
        Equal tables -- bit for bit -- give equal strings, whatever the number of ranks that wrote them."""
        t = self.table()
        parts = []
        for k in self.fields:
            if self.backend == "cuda-ipc":
                torch = self.torch
                w = t[k].view(torch.int64) if self.kind[k] == "f8" else t[k].view(torch.int32).to(torch.int64)
                acc, step = 0, 1 << 26
                for i in range(0, w.numel(), step):  # bounded temporaries for tables of 1e8 rows
                    x = w[i:i + step]
                    idx = torch.arange(i, i + x.numel(), device=x.device, dtype=torch.int64) * 2 + 1
                    acc = (acc + int((x * idx).sum().item())) & 0xFFFFFFFFFFFFFFFF
            else:
                w = t[k].view(np.int64) if self.kind[k] == "f8" else t[k].view(np.int32).astype(np.int64)
                with np.errstate(over="ignore"):
                    acc = int((w * (np.arange(w.size, dtype=np.int64) * 2 + 1)).sum()) & 0xFFFFFFFFFFFFFFFF
            parts.append(f"{acc:016x}")
        return "-".join(parts)

    def close(self):
        """Collective: unmap on the other ranks, then rank 0 frees."""
        import ctypes as C
        if self.backend == "cuda-ipc":
            from . import _cabi
            lib = _cabi.load()
            if self._mapped:
                _cabi.check(lib.pgt_ipc_close(C.c_void_p(self.base)))
            if self.world > 1:
                self.dist.barrier()
            if self.rank == 0 and self.base:
                _cabi.check(lib.pgt_device_free(C.c_void_p(self.base)))
            self.base = None
        else:
            import os
            del self._mm
            if self.world > 1:
                self.dist.barrier()
            if self.rank == 0:
                os.remove(self._path)
