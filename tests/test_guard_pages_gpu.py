"""Kernel memory safety without compute-sanitizer: columns placed against UNMAPPED device pages.

Every input column is put into its own virtual-memory reservation (CUDA VMM: cuMemAddressReserve + cuMemMap)
whose granule before and granule after the mapped range stay unmapped, once with the column starting at the first
mapped byte and once with the column ending at the last mapped byte.  A kernel that reads one 16-byte chunk (an
LDG.128, a bulk-copy superset, a lane past the end of a unit) before or beyond a column faults -- the context dies
and this test fails -- instead of passing silently on whatever lies next to the column in a pooled allocation.
Covered: both level-1 kernels (direct, TMA-tiled), the vectorised genotype kernel, the sliding tile, the per-site
kernel, every statistic, odd lengths; results must also equal those from ordinary allocations, bit for bit."""
import numpy as np
import pytest

from popgenomicstools_b200 import _cabi
from popgenomicstools_b200.sharding import RawDeviceArray

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    try:
        from cuda.bindings import driver as cu
    except Exception:
        pytest.skip("cuda-python driver bindings not available")
    import popgenomicstools_b200 as pgt
    torch.zeros(1, device="cuda")  # primary context
    return pgt, torch, cu


def ok(r):
    assert int(r[0]) == 0, r[0]
    return r[1] if len(r) == 2 else r[1:]


class Guarded:
    """A device buffer of `nbytes` whose neighbours are unmapped: `lo` starts at the first mapped byte, `hi` ends at
    the last mapped byte."""

    def __init__(self, cu, nbytes):
        self.cu = cu
        prop = cu.CUmemAllocationProp()
        prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = 0
        gran = ok(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
        self.size = (max(nbytes, 1) + gran - 1) // gran * gran
        self.total = self.size + 2 * gran
        self.va = int(ok(cu.cuMemAddressReserve(self.total, 0, 0, 0)))
        self.handle = ok(cu.cuMemCreate(self.size, prop, 0))
        ok(cu.cuMemMap(self.va + gran, self.size, 0, self.handle, 0))
        acc = cu.CUmemAccessDesc()
        acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = 0
        acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        ok(cu.cuMemSetAccess(self.va + gran, self.size, [acc], 1))
        self.lo = self.va + gran
        self.hi = self.va + gran + self.size - nbytes

    def free(self):
        cu = self.cu
        gran = (self.total - self.size) // 2
        cu.cuMemUnmap(self.va + gran, self.size)
        cu.cuMemRelease(self.handle)
        cu.cuMemAddressFree(self.va, self.total)


KIND = {"float64": "f8", "int32": "i4", "uint32": "u4", "int8": "i1"}


def guarded_copy(torch, cu, t, at_end, keep):
    g = Guarded(cu, t.numel() * t.element_size())
    keep.append(g)
    ptr = g.hi if at_end else g.lo
    kind = KIND[str(t.dtype).replace("torch.", "")]
    v = torch.as_tensor(RawDeviceArray(ptr, t.numel(), kind), device=t.device)
    v.copy_(t)
    return v


CASES = [("units-direct", 5000, 1000, 0, dict(level1=1)), ("units-tiled", 5000, 1000, 512, dict(level1=2)),
         ("units-tiled-short", 64, 16, 32, dict(level1=2)), ("het-vec", 100000, 100000, 4096, dict(level1=1)),
         ("slide", 1000, 1, 0, {}), ("slide-step7", 777, 7, 0, dict(slide=2)), ("persite", 1, 1, 0, {}), ("scan-mode", 1000, 3, 0, dict(slide=1, level2=2))]


@pytest.mark.parametrize("name,W,S,unit,knobs", CASES)
@pytest.mark.parametrize("n_extra", [0, 1, 7, 13])
def test_no_read_outside_the_columns(env, name, W, S, unit, knobs, n_extra):
    pgt, torch, cu = env
    lengths = [W + 37 * S + n_extra, 200_003 + n_extra, 5]
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    n = int(offs[-1])
    a, b = pgt.synth_fst(41, 0, n)
    g = pgt.synth_het(41, 0, n)
    f1, f2, n1, n2 = pgt.synth_dxy(41, 0, n)
    pos = pgt.synth_pos(41, 0, n, offs, 2)
    cols = dict(pos=pos, a=a, b=b, geno=g, f1=f1, f2=f2, n1=n1, n2=n2)
    plan = pgt.WindowPlan(offs, W, S, unit_sites=unit)
    try:
        for k, v in knobs.items():
            pgt.tune(k, v)
        ref = pgt.scan(plan, _cabi.PGT_STAT_FUSED, cols, minind=5)
        torch.cuda.synchronize()
        for at_end in (False, True):
            keep = []
            gc = {k: guarded_copy(torch, cu, v, at_end, keep) for k, v in cols.items()}
            torch.cuda.synchronize()
            for stat, need in ((_cabi.PGT_STAT_FUSED, list(cols)), (_cabi.PGT_STAT_FST, ["pos", "a", "b"]), (_cabi.PGT_STAT_HET, ["pos", "geno"]),
                               (_cabi.PGT_STAT_DXY, ["pos", "f1", "f2", "n1", "n2"])):
                res = pgt.scan(plan, stat, {k: gc[k] for k in need}, minind=5)
                torch.cuda.synchronize()  # an out-of-bounds access surfaces here as an illegal-address error
                for k, v in res.items():
                    got, want = v.cpu().numpy(), ref[k].cpu().numpy()
                    if k == "dxy_global":  # the one output whose summation order may follow the kernel: counts exact
                        assert got[1] == want[1] and got[2] == want[2] and abs(got[0] - want[0]) <= 1e-12 * abs(want[0]), (name, at_end, stat)
                    else:
                        assert got.tobytes() == want.tobytes(), (name, at_end, stat, k)
            del gc
            for gbuf in keep:
                gbuf.free()
    finally:
        for k in knobs:
            pgt.tune(k, 0)
