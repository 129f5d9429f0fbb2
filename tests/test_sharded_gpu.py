"""pgt_scan_sharded / pgt_scan_extreme_sharded / PGT_DEVICES: one host column set, several GPUs of ONE process,
one host thread per GPU, every shard's rows copied straight into the caller's table at its window offset (no
gather).  The reference is a single process (/root/reference/fstWindow.cpp:158-177): the drop-in must give the
same stdout whatever the device list.  On a single-GPU box the lists repeat device 0 (the shards then run side by
side on it, each on its own stream); with more GPUs visible the same tests also spread over all of them."""
import numpy as np
import pytest

import cli_util as U
import oracle_lib as O
import textfmt as T
from popgenomicstools_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pgt():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import popgenomicstools_b200 as m
    return m


def offsets(lengths):
    return np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)


def device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0], [0, 0, 0], [0] * 7]
    if n >= 2:
        lists += [list(range(n)), list(range(n))[::-1]]
    return lists


@pytest.mark.parametrize("W,S,unit", [(5000, 1000, 0), (1000, 100, 512), (1000, 1, 0), (1, 1, 0), (300, 7, 32)])
def test_fused_table_is_identical_for_any_device_list(pgt, W, S, unit):
    lengths = [W + 40 * S, 250_000, 3, W // 2 + 1, 90_000]
    offs = offsets(lengths)
    n = int(offs[-1])
    a, b = O.synth_fst(21, 0, n)
    g = O.synth_het(21, 0, n)
    f1, f2, n1, n2 = O.synth_dxy(21, 0, n)
    pos = O.synth_pos(21, offs, 2)
    cols = dict(pos=pos, a=a, b=b, geno=g, f1=f1, f2=f2, n1=n1, n2=n2)
    plan = pgt.WindowPlan(offs, W, S, unit_sites=unit)
    ref = pgt.scan(plan, _cabi.PGT_STAT_FUSED, cols, minind=4)  # the ordinary single-device host-memory scan
    for devs in device_lists():
        res = pgt.scan_sharded(plan, _cabi.PGT_STAT_FUSED, cols, devs, minind=4)
        for k, v in ref.items():
            if k == "dxy_global":
                assert res[k][1] == v[1] and res[k][2] == v[2] and abs(res[k][0] - v[0]) <= 1e-12 * abs(v[0]), (devs, res[k], v)
            else:
                assert res[k].tobytes() == v.tobytes(), (devs, k)
    # and it is the oracle's table
    rh = O.het(T.expand_chr(lengths), pos, g, W, S)
    assert np.array_equal(ref["nhet"], rh["nhet"]) and np.array_equal(ref["start_pos"], rh["start"])


def test_bp_mode_and_empty_shards(pgt):
    """dxyWindow -fixedsite 0 through the sharded call (columns address the whole axis; every shard searches its
    own slabs), more shards than windows, and a plan without any window (shard 0 alone owns the global line)."""
    nsites = [30_000, 12_000, 700]
    chr_len = [300_000, 150_000, 9_000]
    soff = offsets(nsites)
    n = int(soff[-1])
    f1, f2, n1, n2 = O.synth_dxy(22, 0, n)
    pos = O.synth_pos(22, soff, 9)
    cols = dict(pos=pos, f1=f1, f2=f2, n1=n1, n2=n2)
    plan = pgt.WindowPlan(offsets(chr_len), 20000, 5000, mode="bp")
    ref = pgt.scan(plan, _cabi.PGT_STAT_DXY, cols, minind=5, site_offsets=soff)
    for devs in device_lists():
        res = pgt.scan_sharded(plan, _cabi.PGT_STAT_DXY, cols, devs, minind=5, site_offsets=soff)
        for k, v in ref.items():
            if k == "dxy_global":
                assert res[k][1] == v[1] and res[k][2] == v[2] and abs(res[k][0] - v[0]) <= 1e-12 * abs(v[0])
            else:
                assert res[k].tobytes() == v.tobytes(), (devs, k)
    few = pgt.WindowPlan(offsets([1200]), 1000, 100)  # 3 windows, up to 7 shards
    a, b = O.synth_fst(23, 0, 1200)
    p1 = O.synth_pos(23, offsets([1200]), 1)
    r1 = pgt.scan(few, _cabi.PGT_STAT_FST, dict(pos=p1, a=a, b=b))
    for devs in device_lists():
        r = pgt.scan_sharded(few, _cabi.PGT_STAT_FST, dict(pos=p1, a=a, b=b), devs)
        for k, v in r1.items():
            assert r[k].tobytes() == v.tobytes(), (devs, k)
    none = pgt.WindowPlan(offsets([50]), 1000, 100)  # EOF partial of 50 <= W - S sites: no window, global line only
    assert none.num_windows == 0
    d1, d2, m1, m2 = O.synth_dxy(24, 0, 50)
    c0 = dict(pos=np.arange(1, 51, dtype=np.uint32), f1=d1, f2=d2, n1=m1, n2=m2)
    g1 = pgt.scan(none, _cabi.PGT_STAT_DXY, c0, minind=5)["dxy_global"]
    for devs in device_lists():
        g = pgt.scan_sharded(none, _cabi.PGT_STAT_DXY, c0, devs, minind=5)["dxy_global"]
        assert g[1] == g1[1] and g[2] == g1[2] and g[1] + g[2] == 50 and abs(g[0] - g1[0]) <= 1e-12 * abs(g1[0])
    with pytest.raises(pgt.PgtError):
        pgt.scan_sharded(few, _cabi.PGT_STAT_FST, dict(pos=p1, a=a, b=b), [99])


def test_extreme_scan_sharded(pgt):
    xoff = offsets([40_000, 25_000, 9_000])
    n = int(xoff[-1])
    pos = O.synth_pos(25, xoff, 41)
    import torch
    score = pgt.synth_score(25, 0, n).cpu().numpy()
    xplan = pgt.ExtremePlan(pos, xoff, 10000)
    ref = pgt.scan_extreme(xplan, _cabi.PGT_XSTAT_IHS, 2.0, pos, score)
    for devs in device_lists():
        res = pgt.scan_extreme_sharded(xplan, _cabi.PGT_XSTAT_IHS, 2.0, pos, score, devs)
        for k, v in ref.items():
            assert res[k].tobytes() == v.tobytes(), (devs, k)


def test_clis_print_the_same_rows_with_pgt_devices(pgt, tmp_path):
    """PGT_DEVICES on all five tools: byte-identical stdout (and stderr: dxyWindow's global line) for 1, 2 and 5
    shards.  CUDA_VISIBLE_DEVICES is set here, so the list is taken literally and may repeat a device."""
    import torch
    names = ["chr1", "chr2", "chrUn"]
    lengths = [60_000, 41_000, 900]
    offs = offsets(lengths)
    O.write_text("fst", str(tmp_path / "s.fst"), names, offs, seed=31, density=3)
    O.write_text("het", str(tmp_path / "s.het"), names, offs, seed=31, density=3)
    rng = np.random.default_rng(31)
    pos = np.concatenate([np.cumsum(rng.integers(1, 9, size=L)) for L in lengths])
    f1, f2 = rng.integers(0, 1000001, size=int(offs[-1])), rng.integers(0, 1000001, size=int(offs[-1]))
    n1, n2 = rng.integers(0, 30, size=int(offs[-1])), rng.integers(0, 30, size=int(offs[-1]))
    (tmp_path / "p1.mafs").write_text(T.maf_text(names, lengths, pos, f1, n1))
    (tmp_path / "p2.mafs").write_text(T.maf_text(names, lengths, pos, f2, n2))
    (tmp_path / "sizes.txt").write_text("".join(f"{nm}\t{int(pos[offs[i + 1] - 1]) + 77}\n" for i, nm in enumerate(names)))
    v = rng.integers(-4000000, 4000001, size=int(offs[-1]))
    (tmp_path / "i.norm").write_text(T.ihs_text(names, lengths, pos * 37, v))
    (tmp_path / "x.norm").write_text(T.xpehh_text(names, lengths, pos * 37, v))
    vis = ",".join(str(i) for i in range(torch.cuda.device_count()))
    lists = ["0,0", "0,0,0,0,0"] + ([vis] if torch.cuda.device_count() > 1 else [])
    runs = [("fstWindow", ["s.fst", 5000, 1000]), ("fstWindow", ["s.fst", 1000, 1]), ("hetWindow", ["s.het", 20000, 20000]),
            ("hetWindow", ["s.het"]),
            ("dxyWindow", ["-winsize", 20000, "-stepsize", 5000, "-minind", 5, "-sizefile", "sizes.txt", "p1.mafs", "p2.mafs"]),
            ("dxyWindow", ["-winsize", 700, "-stepsize", 50, "-fixedsite", 1, "p1.mafs", "p2.mafs"]),
            ("ihsWindow", ["i.norm", "-winsize", 100000]), ("xpehhWindow", ["x.norm", "-2", "-winsize", 100000])]
    for tool, args in runs:
        one = U.run(U.ours(tool), args, cwd=str(tmp_path), env={"CUDA_VISIBLE_DEVICES": vis})
        assert one[0] == 0 and one[1], (tool, args, one[2])
        for lst in lists:
            many = U.run(U.ours(tool), args, cwd=str(tmp_path), env={"CUDA_VISIBLE_DEVICES": vis, "PGT_DEVICES": lst})
            assert many == one, (tool, args, lst, many[2][:300])
