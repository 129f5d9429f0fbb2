"""Which kernels a scan runs is host logic (pgt_plan_scan_path: no CUDA call) and fixes the summation order, so it must
be a function of (W, S, unit, statistic) only.  Also the shard arithmetic the one-process multi-GPU call relies on:
pgt_scan_sharded_workspace_bytes for every shard, including shards without windows."""
import numpy as np
import pytest

import popgenomicstools_b200 as pgt
from popgenomicstools_b200 import _cabi

STATS = (_cabi.PGT_STAT_FST, _cabi.PGT_STAT_HET, _cabi.PGT_STAT_DXY, _cabi.PGT_STAT_FUSED)


def offsets(lengths):
    return np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)


def test_scan_path_rule():
    """sliding tile iff no piece of a step reaches 32 sites (max(W % S, S - W % S) < 32), a window has more than 32 units
    and at most 1048 sites (the fused statistic's step must fit shared memory; ONE rule for all statistics, so that the
    fused scan equals the three single scans bit for bit); W = S = 1 is the per-site map; everything else units."""
    want = {(1000, 1): "slide", (1000, 7): "slide", (1040, 16): "slide", (1048, 1): "slide", (1049, 1): "units", (1000, 31): "slide",
            (1000, 32): "slide", (1010, 40): "slide", (255, 1): "slide", (64, 1): "slide", (100, 3): "slide", (33, 1): "slide",
            (32, 1): "units", (1000, 40): "units", (1024, 32): "units", (1000, 100): "units", (2048, 1): "units",
            (50000, 10000): "units", (100000, 100000): "units", (1, 1): "persite", (2, 1): "units", (2, 2): "units"}
    for lengths in ([50000, 7000], [3], [10 ** 9, 2 * 10 ** 9]):  # the input size and the contig layout do not enter
        offs = offsets(lengths)
        for (W, S), path in want.items():
            for unit in (0, 32, 512, 4096):
                for stat in STATS:
                    assert pgt.WindowPlan(offs, W, S, unit_sites=unit).scan_path(stat) == path, (lengths, W, S, unit, stat)
    assert pgt.WindowPlan(offsets([50000]), 1000, 1, mode="bp").scan_path(_cabi.PGT_STAT_DXY) == "units"
    assert pgt.WindowPlan(offsets([50000]), 1, 1, mode="bp").scan_path(_cabi.PGT_STAT_DXY) == "units"
    try:
        pgt.tune("slide", 1)
        assert pgt.WindowPlan(offsets([50000]), 1000, 1).scan_path(_cabi.PGT_STAT_FST) == "units"
        pgt.tune("slide", 2)  # forced: every site-mode geometry whose step fits
        assert pgt.WindowPlan(offsets([50000]), 5, 2).scan_path(_cabi.PGT_STAT_FUSED) == "slide"
        assert pgt.WindowPlan(offsets([50000]), 50000, 10000).scan_path(_cabi.PGT_STAT_FST) == "units"
        pgt.tune("level2", 1)
        pgt.tune("slide", 0)
        assert pgt.WindowPlan(offsets([50000]), 1, 1).scan_path(_cabi.PGT_STAT_HET) == "units"
    finally:
        pgt.tune("slide", 0)
        pgt.tune("level2", 0)
    with pytest.raises(pgt.PgtError):
        pgt.WindowPlan(offsets([5]), 3, 1).scan_path(99)


def test_workspace_queries_need_no_device():
    """Sizing is pure host arithmetic: bounded for the unit-free paths in host mode, and defined for every shard."""
    lib = _cabi.load()
    big = pgt.WindowPlan(offsets([2_000_000_000, 1_000_000_000]), 1000, 1)
    assert big.workspace_bytes(_cabi.PGT_STAT_FUSED, _cabi.PGT_MEM_HOST) < (3 << 29)
    assert big.workspace_bytes(_cabi.PGT_STAT_FST, _cabi.PGT_MEM_DEVICE) < (1 << 20)  # no unit array at all
    units = pgt.WindowPlan(offsets([2_000_000_000, 1_000_000_000]), 50000, 10000, unit_sites=512)
    w = units.workspace_bytes(_cabi.PGT_STAT_FST, _cabi.PGT_MEM_DEVICE)
    assert 16 * units.num_units <= w < 2 * 16 * units.num_units + (1 << 22)
    few = pgt.WindowPlan(offsets([1200]), 1000, 100)  # 3 windows
    sizes = [int(lib.pgt_scan_sharded_workspace_bytes(few.handle, _cabi.PGT_STAT_FST, r, 7)) for r in range(7)]
    assert all(s >= 256 for s in sizes) and max(sizes) < (1 << 30)
    assert int(lib.pgt_scan_sharded_workspace_bytes(few.handle, _cabi.PGT_STAT_FST, 7, 7)) == 0  # shard out of range
