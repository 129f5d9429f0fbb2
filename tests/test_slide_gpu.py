"""Fine steps under long windows (W = 1000, S = 1 and kin): the sliding-tile kernel k_slide forms the
windows straight from the sites in shared memory (block prefix / suffix sums on W-aligned blocks of the
segment, window = SUF[first] + PRE[last]) instead of going through a unit array.  What the reference
does per window is a re-sum of its buffer (/root/reference/fstWindow.cpp:80-99, hetWindow.cpp:77-97,
dxyWindow.cpp:179-201); the oracle restates that.  Checked here: against the oracle (exact for
membership, positions, labels and the integer statistics; stated tolerance for the sums), device and
host memory, shards bit-identical, unaligned columns, the global line, and the path selection rule."""
import numpy as np
import pytest

import oracle_lib as O
import parity as P
import textfmt as T
from popgenomicstools_b200 import _cabi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pgt():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import popgenomicstools_b200 as m
    yield m
    m.tune("slide", 0)


def npy(out):
    import torch
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}


def offsets(lengths):
    return np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)


def columns(pgt, seed, offs, density=1):
    n = int(offs[-1])
    a, b = pgt.synth_fst(seed, 0, n)
    g = pgt.synth_het(seed, 0, n)
    f1, f2, n1, n2 = pgt.synth_dxy(seed, 0, n)
    pos = pgt.synth_pos(seed, 0, n, offs, density)
    d = dict(pos=pos, a=a, b=b, geno=g, f1=f1, f2=f2, n1=n1, n2=n2)
    return d, {k: v.cpu().numpy() for k, v in d.items()}


def oracle_all(lengths, h, W, S, minind):
    chr_id = T.expand_chr(lengths)
    rf = O.fst(chr_id, h["pos"], h["a"], h["b"], W, S)
    ra = O.fst(chr_id, h["pos"], np.abs(h["a"]), np.abs(h["b"]), W, S)
    rh = O.het(chr_id, h["pos"], h["geno"], W, S)
    rd = O.dxy(chr_id, h["pos"], h["f1"], h["f2"], h["n1"], h["n2"], minind, W, S, 1)
    return rf, ra, rh, rd


def check_fused(res, refs, tag):
    rf, ra, rh, rd = refs
    P.assert_exact(res["label"], rf["label"], tag + " label")
    P.assert_exact(res["start_pos"], rf["start"], tag + " start")
    P.assert_exact(res["end_pos"], rf["end"], tag + " end")
    P.assert_exact(res["mid_pos"], rf["mid"], tag + " mid")
    P.assert_exact(res["nsites"], rf["n"], tag + " nsites")
    P.assert_sum_close(res["sum_a"], rf["asum"], ra["asum"], tag + " sum_a")
    P.assert_sum_close(res["sum_b"], rf["bsum"], ra["bsum"], tag + " sum_b")
    fst = np.where(res["sum_b"] != 0, res["sum_a"] / np.where(res["sum_b"] != 0, res["sum_b"], 1.0), 0.0)
    P.assert_exact(res["fst"], fst, tag + " fst = sum_a / sum_b (one IEEE divide)")
    P.assert_exact(res["nhet"], rh["nhet"], tag + " nhet")
    P.assert_exact(res["nonmissing"], rh["nonmissing"], tag + " nonmissing")
    P.assert_exact(res["het"], rh["h"], tag + " het")
    P.assert_exact(res["neffective"], rd["neff"], tag + " neff")
    P.assert_exact(res["nskip"], rd["nskip"], tag + " nskip")
    P.assert_sum_close(res["dxy"], rd["dxy"], rd["dxy"], tag + " dxy")
    g, rg = res["dxy_global"], rd["global"]
    assert g[1] == rg[1] and g[2] == rg[2], (tag, g, rg)
    assert abs(g[0] - rg[0]) <= 1e-9 * abs(rg[0]) + 1e-300, (tag, g, rg)


SHAPES = [(1, 1), (2, 1), (5, 2), (31, 1), (64, 64), (255, 7), (256, 1), (257, 3), (300, 299), (777, 13), (1000, 1), (1000, 7),
          (1024, 32), (1047, 5), (1048, 1), (1048, 1048), (129, 1), (192, 5), (500, 3), (700, 1)]


@pytest.mark.parametrize("W,S", SHAPES)
def test_sliding_tile_forced_on_every_shape_matches_the_oracle(pgt, W, S):
    """pgt_tune slide=2 sends every site-mode geometry with W <= 1048 through k_slide: blocks shorter and
    longer than the consumer team, steps that do not divide the window, windows straddling two blocks, the
    cross-contig carry (first contig), trailing partial windows, contigs shorter than a window, EOF drop."""
    lengths = [W + 4 * S, 6011, 3, W // 2 + 1, 2 * W + 1, 4000, max(1, W - S)]
    offs = offsets(lengths)
    minind = 5
    try:
        pgt.tune("slide", 2)
        d, h = columns(pgt, 11, offs, density=3)
        refs = oracle_all(lengths, h, W, S, minind)
        plan = pgt.WindowPlan(offs, W, S)
        assert plan.scan_path(_cabi.PGT_STAT_FUSED) == "slide"
        assert plan.num_windows == len(refs[0]["n"])
        res = npy(pgt.fused_window(plan, d["pos"], d["a"], d["b"], d["geno"], d["f1"], d["f2"], d["n1"], d["n2"], minind=minind))
        check_fused(res, refs, f"W={W} S={S} device")
        # each single-statistic scan gives the fused table's columns bit for bit
        for one in (pgt.fst_window(plan, d["pos"], d["a"], d["b"]), pgt.het_window(plan, d["pos"], d["geno"]),
                    pgt.dxy_window(plan, d["pos"], d["f1"], d["f2"], d["n1"], d["n2"], minind=minind)):
            for k, v in npy(one).items():
                if k != "dxy_global":
                    assert v.tobytes() == res[k].tobytes(), (W, S, k)
        # host-memory mode (slabs of windows through the staging buffers): the same table
        hres = pgt.fused_window(plan, h["pos"], h["a"], h["b"], h["geno"], h["f1"], h["f2"], h["n1"], h["n2"], minind=minind)
        for k in res:
            if k == "dxy_global":
                assert hres[k][1] == res[k][1] and hres[k][2] == res[k][2] and abs(hres[k][0] - res[k][0]) <= 1e-12 * abs(res[k][0])
            else:
                assert hres[k].tobytes() == res[k].tobytes(), (W, S, k)
        # shards: the unsharded table bit for bit, for any shard count
        for nsh in (2, 5):
            parts = []
            for r in range(nsh):
                wl, wh, sl, sh = plan.shard(r, nsh)
                if wh > wl:
                    parts.append(npy(pgt.fused_window(plan, *[d[k][sl:sh] for k in ("pos", "a", "b", "geno", "f1", "f2", "n1", "n2")],
                                                      minind=minind, window_range=(wl, wh), site_origin=sl)))
            for k in res:
                if k != "dxy_global":
                    assert np.concatenate([p[k] for p in parts]).tobytes() == res[k].tobytes(), (W, S, nsh, k)
            gl = np.sum([p["dxy_global"] for p in parts], axis=0)
            assert gl[1] == res["dxy_global"][1] and gl[2] == res["dxy_global"][2]
    finally:
        pgt.tune("slide", 0)


@pytest.mark.parametrize("lengths,W,S", [([20000], 1000, 1), ([20000], 300, 7), ([9000, 400, 5000, 2, 7000], 1000, 1), ([6000, 6001], 1048, 1048)])
def test_global_line_taken_inside_the_sliding_tile(pgt, lengths, W, S):
    """dxyWindow's global line (dxyWindow.cpp:382-385) over device-resident columns: the sliding tile adds up the
    sites its runs own on the way -- one launch less, the dxy columns read once -- when those are exactly the line's
    sites, else a pass over the columns takes it.  Either way: the oracle's counts exactly, its sum to 1e-9; the
    two ways agree to 1e-12; shards add up."""
    offs = offsets(lengths)
    d, h = columns(pgt, 23, offs, density=2)
    rd = O.dxy(T.expand_chr(lengths), h["pos"], h["f1"], h["f2"], h["n1"], h["n2"], 5, W, S, 1)
    try:
        pgt.tune("slide", 2)
        plan = pgt.WindowPlan(offs, W, S)
        got = {}
        for knob in (0, 1):
            pgt.tune("slideglobal", knob)
            for name, fn in (("dxy", lambda: pgt.dxy_window(plan, d["pos"], d["f1"], d["f2"], d["n1"], d["n2"], minind=5)),
                             ("fused", lambda: pgt.fused_window(plan, d["pos"], d["a"], d["b"], d["geno"], d["f1"], d["f2"], d["n1"], d["n2"], minind=5))):
                fn()
                n0 = pgt.kernel_launch_count()
                res = npy(fn())
                got[(name, knob)] = (res, pgt.kernel_launch_count() - n0)
                g = res["dxy_global"]
                assert g[1] == rd["global"][1] and g[2] == rd["global"][2], (name, knob, g, rd["global"])
                assert abs(g[0] - rd["global"][0]) <= 1e-9 * abs(rd["global"][0]) + 1e-300
        for name in ("dxy", "fused"):
            (a, la), (b, lb) = got[(name, 0)], got[(name, 1)]
            assert abs(a["dxy_global"][0] - b["dxy_global"][0]) <= 1e-12 * abs(b["dxy_global"][0]) + 1e-300
            for k in a:
                if k != "dxy_global":
                    assert a[k].tobytes() == b[k].tobytes(), (name, k)
            if len(lengths) == 1:  # one contig longer than a window: every site lies in a window, the tile takes the line
                assert la == lb - 1, (name, la, lb)
        pgt.tune("slideglobal", 0)
        for nsh in (2, 3):
            tot = np.zeros(3)
            for r in range(nsh):
                wl, wh, sl, sh = plan.shard(r, nsh)
                tot += npy(pgt.dxy_window(plan, d["pos"][sl:sh], d["f1"][sl:sh], d["f2"][sl:sh], d["n1"][sl:sh], d["n2"][sl:sh], minind=5,
                                          window_range=(wl, wh), site_origin=sl))["dxy_global"]
            assert tot[1] == rd["global"][1] and tot[2] == rd["global"][2], (nsh, tot, rd["global"])
    finally:
        pgt.tune("slide", 0)
        pgt.tune("slideglobal", 0)


def test_path_selection_is_a_function_of_the_geometry_only(pgt):
    """auto: sliding tile iff no piece of a step reaches 32 sites (max(W % S, S - W % S) < 32) under windows of more
    than 32 units and at most 1048 sites (the fused statistic's step must fit shared memory; one rule for all statistics)."""
    offs = offsets([50000, 7000])
    want = {(1000, 1): "slide", (1000, 7): "slide", (1040, 16): "slide", (1048, 1): "slide", (1049, 1): "units", (1000, 31): "slide",
            (1000, 32): "slide", (1010, 40): "slide",  # pieces of 10 and 30 sites
            (255, 1): "slide", (64, 1): "slide", (100, 3): "slide", (32, 1): "units",  # 32 one-site units: thread-per-window level 2
            (1000, 40): "units", (1024, 32): "units", (1000, 100): "units", (2048, 1): "units",
            (50000, 10000): "units", (1, 1): "persite"}
    for (W, S), path in want.items():
        for stat in (_cabi.PGT_STAT_FST, _cabi.PGT_STAT_HET, _cabi.PGT_STAT_DXY, _cabi.PGT_STAT_FUSED):
            assert pgt.WindowPlan(offs, W, S).scan_path(stat) == path, (W, S, stat)
    assert pgt.WindowPlan(offs, 1000, 1, mode="bp").scan_path(_cabi.PGT_STAT_DXY) == "units"
    try:
        pgt.tune("slide", 1)
        assert pgt.WindowPlan(offs, 1000, 1).scan_path(_cabi.PGT_STAT_FST) == "units"
    finally:
        pgt.tune("slide", 0)


@pytest.mark.parametrize("W,S", [(1000, 1), (1000, 7), (300, 2)])
def test_auto_path_against_oracle_and_against_the_unit_path(pgt, W, S):
    """Default knobs at the stress shape: the sliding tile is chosen, matches the oracle, agrees with the
    two-level unit path (slide=1) within the summation tolerance and exactly on everything integer, and
    tolerates columns that are only element-aligned (views starting at odd elements)."""
    import torch
    lengths = [W + 250 * S, 131071, 17, 64000]
    offs = offsets(lengths)
    d, h = columns(pgt, 12, offs)
    refs = oracle_all(lengths, h, W, S, 3)
    plan = pgt.WindowPlan(offs, W, S)
    assert plan.scan_path(_cabi.PGT_STAT_FUSED) == "slide"
    args = [d[k] for k in ("pos", "a", "b", "geno", "f1", "f2", "n1", "n2")]
    res = npy(pgt.fused_window(plan, *args, minind=3))
    check_fused(res, refs, f"auto W={W} S={S}")
    # unaligned views: element 0 of every column sits 1 (or 3) elements into its allocation
    for shift in (1, 3):
        sh_args = []
        for x in args:
            buf = torch.empty(x.numel() + 8, dtype=x.dtype, device=x.device)
            buf[shift:shift + x.numel()].copy_(x)
            sh_args.append(buf[shift:shift + x.numel()])
        r2 = npy(pgt.fused_window(plan, *sh_args, minind=3))
        for k in res:
            assert r2[k].tobytes() == res[k].tobytes(), (shift, k)
    try:
        pgt.tune("slide", 1)
        assert plan.scan_path(_cabi.PGT_STAT_FUSED) == "units"
        ru = npy(pgt.fused_window(plan, *args, minind=3))
    finally:
        pgt.tune("slide", 0)
    check_fused(ru, refs, f"unit path W={W} S={S}")
    for k in ("label", "start_pos", "end_pos", "mid_pos", "nsites", "nhet", "nonmissing", "het", "neffective", "nskip"):
        assert ru[k].tobytes() == res[k].tobytes(), k


def test_many_short_contigs_through_the_sliding_tile(pgt):
    """Chunks that span hundreds of segments (contigs shorter than the window: one partial window each)."""
    rng = np.random.default_rng(5)
    W, S = 1000, 3
    lengths = [int(x) for x in rng.integers(1, 2500, size=700)]
    offs = offsets(lengths)
    d, h = columns(pgt, 13, offs)
    refs = oracle_all(lengths, h, W, S, 5)
    plan = pgt.WindowPlan(offs, W, S)
    assert plan.scan_path(_cabi.PGT_STAT_FUSED) == "slide"
    res = npy(pgt.fused_window(plan, *[d[k] for k in ("pos", "a", "b", "geno", "f1", "f2", "n1", "n2")], minind=5))
    check_fused(res, refs, "700 contigs")
    hres = pgt.fst_window(plan, h["pos"], h["a"], h["b"])
    for k in ("sum_a", "sum_b", "fst", "start_pos", "end_pos", "label", "nsites"):
        assert hres[k].tobytes() == res[k].tobytes(), k


def test_default_arguments_in_host_mode_need_no_unit_array(pgt):
    """W = S = 1 (the tools' default argv) in PGT_MEM_HOST runs the per-site kernel slab by slab: the
    workspace stays ~1 GB whatever the input size (it used to hold one 16-byte unit and 96 bytes of staged outputs
    per site)."""
    n = 6_000_000
    offs = offsets([n - 5, 5])
    plan = pgt.WindowPlan(offs, 1, 1)
    assert plan.scan_path(_cabi.PGT_STAT_FST) == "persite"
    # bounded workspace: two slabs of columns and two slab-sized output tables, whatever the genome size
    assert plan.workspace_bytes(_cabi.PGT_STAT_FST, _cabi.PGT_MEM_HOST) < (1 << 30)
    big = pgt.WindowPlan(offsets([2_000_000_000, 1_000_000_000]), 1, 1)
    assert big.workspace_bytes(_cabi.PGT_STAT_FUSED, _cabi.PGT_MEM_HOST) < (3 << 29)
    big = pgt.WindowPlan(offsets([2_000_000_000, 1_000_000_000]), 1000, 1)
    assert big.workspace_bytes(_cabi.PGT_STAT_FUSED, _cabi.PGT_MEM_HOST) < (3 << 29)
    a, b = O.synth_fst(7, 0, n)
    pos = O.synth_pos(7, offs, 1)
    res = pgt.fst_window(plan, pos, a, b)
    assert np.array_equal(res["sum_a"], 0.0 + a) and np.array_equal(res["sum_b"], 0.0 + b)
    assert np.array_equal(res["start_pos"], pos) and np.array_equal(res["end_pos"], pos) and np.all(res["nsites"] == 1)
    f1, f2, n1, n2 = O.synth_dxy(7, 0, n)
    rd = pgt.dxy_window(plan, pos, f1, f2, n1, n2, minind=5)
    ref = O.dxy(T.expand_chr([n - 5, 5]), pos, f1, f2, n1, n2, 5, 1, 1, 1)
    P.assert_exact(rd["neffective"], ref["neff"], "neff")
    P.assert_exact(rd["dxy"], ref["dxy"], "per-site dxy (bit-exact: no FMA, dxyWindow.cpp:381)")
    assert rd["dxy_global"][1] == ref["global"][1] and rd["dxy_global"][2] == ref["global"][2]
    assert abs(rd["dxy_global"][0] - ref["global"][0]) <= 1e-9 * ref["global"][0]


def test_host_mode_over_several_slabs_equals_device_mode(pgt):
    """PGT_MEM_HOST stages 4 Mi-site slabs; sliding-tile scans go slab by slab over WINDOWS (neighbouring slabs share
    W - S sites, the rows of a slab are copied back while the next one uploads).  9e6 sites = three slabs, several
    contigs with the slab cuts falling inside them: the table must equal the device-mode table bit for bit, also
    through pgt_scan_sharded with three shards."""
    lengths = [5_000_003, 2_999_999, 17, 1_000_000]
    offs = offsets(lengths)
    d, h = columns(pgt, 14, offs)
    for W, S in ((1000, 3), (200, 1)):
        plan = pgt.WindowPlan(offs, W, S)
        assert plan.scan_path(_cabi.PGT_STAT_FST) == "slide"
        dev = npy(pgt.fst_window(plan, d["pos"], d["a"], d["b"]))
        host = pgt.fst_window(plan, h["pos"], h["a"], h["b"])
        for k in dev:
            assert host[k].tobytes() == dev[k].tobytes(), (W, S, k)
        plan_h = pgt.WindowPlan(offs, W, S)  # an unbound plan for the one-process multi-device call
        sh = pgt.scan_sharded(plan_h, _cabi.PGT_STAT_FST, dict(pos=h["pos"], a=h["a"], b=h["b"]), [0, 0, 0])
        for k in dev:
            assert sh[k].tobytes() == dev[k].tobytes(), (W, S, k)
    # dxy with the global line over the slabs (per-slab partial lines added in order)
    plan = pgt.WindowPlan(offs, 1000, 3)
    dd = npy(pgt.dxy_window(plan, d["pos"], d["f1"], d["f2"], d["n1"], d["n2"], minind=5))
    hd = pgt.dxy_window(plan, h["pos"], h["f1"], h["f2"], h["n1"], h["n2"], minind=5)
    for k in dd:
        if k == "dxy_global":
            assert hd[k][1] == dd[k][1] and hd[k][2] == dd[k][2] and abs(hd[k][0] - dd[k][0]) <= 1e-12 * dd[k][0]
        else:
            assert hd[k].tobytes() == dd[k].tobytes(), k
