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
