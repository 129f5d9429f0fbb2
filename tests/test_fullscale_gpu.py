"""Parity at BASELINE.json's full sizes.

C2 (dxy, 10 Mb contig, 20 kb / 5 kb bp windows, 1e6 and 1e7 sites) and C3 (het, 1e8 sites,
single-site and 100 kb windows) are small enough for the oracle to check EVERY window.
C4 (fst, 3e9 sites / 24 contigs, 50000 / 10000) is checked through size-independent properties:
sampled windows against an exactly rounded sum of their sites (math.fsum), tiling identities
between window sets, shard invariance, and the two level-1 kernels against each other."""
import math

import numpy as np
import pytest

import oracle_lib as O
import parity as P
import textfmt as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pgt():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import popgenomicstools_b200 as m
    return m


def npy(out):
    import torch
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("density", [10, 1])
def test_c2_dxy_10mb_contig_all_windows(pgt, density):
    L, W, S = 10_000_000, 20000, 5000
    n = L // density
    soff = np.array([0, n], np.uint64)
    f1, f2, n1, n2 = pgt.synth_dxy(2, 0, n)
    pos = pgt.synth_pos(2, 0, n, soff, density)
    plan = pgt.WindowPlan([0, L], W, S, mode="bp")
    assert plan.num_windows == 1997
    res = npy(pgt.dxy_window(plan, pos, f1, f2, n1, n2, minind=5, site_offsets=soff))
    ref = O.dxy(np.zeros(n, np.uint32), pos.cpu().numpy(), f1.cpu().numpy(), f2.cpu().numpy(), n1.cpu().numpy(),
                n2.cpu().numpy(), 5, W, S, 0, 0, [L])
    P.assert_exact(res["start_pos"].astype(np.int64), ref["start"], "start")
    P.assert_exact(res["end_pos"].astype(np.int64), ref["end"], "end")
    P.assert_exact(res["neffective"], ref["neff"], "neff")
    P.assert_exact(res["nskip"], ref["nskip"], "nskip")
    P.assert_sum_close(res["dxy"], ref["dxy"], ref["dxy"], "dxy")
    assert (res["start_pos"][-1], res["end_pos"][-1]) == (9980001, 10000000)
    assert res["dxy_global"][1] == ref["global"][1] and res["dxy_global"][2] == ref["global"][2]
    assert abs(res["dxy_global"][0] - ref["global"][0]) <= 1e-9 * ref["global"][0]
    # -fixedsite 1 on the same sites (197 windows at 1e6 sites)
    plan_s = pgt.WindowPlan(soff, W, S)
    res = npy(pgt.dxy_window(plan_s, pos, f1, f2, n1, n2, minind=5))
    ref = O.dxy(np.zeros(n, np.uint32), pos.cpu().numpy(), f1.cpu().numpy(), f2.cpu().numpy(), n1.cpu().numpy(),
                n2.cpu().numpy(), 5, W, S, 1)
    if density == 10:
        assert plan_s.num_windows == 197
    P.assert_exact(res["neffective"], ref["neff"], "neff")
    P.assert_sum_close(res["dxy"], ref["dxy"], ref["dxy"], "dxy")


@pytest.mark.parametrize("W,S,unit", [(1, 1, 0), (100000, 100000, 4096), (100000, 20000, 0)])
def test_c3_het_100mb_chromosome_all_windows(pgt, W, S, unit):
    n = 100_000_000
    offs = np.array([0, n], np.uint64)
    g = pgt.synth_het(3, 0, n)
    pos = pgt.synth_pos(3, 0, n, offs, 1)
    plan = pgt.WindowPlan(offs, W, S, unit_sites=unit)
    res = npy(pgt.het_window(plan, pos, g))
    ref = O.het(np.zeros(n, np.uint32), pos.cpu().numpy(), g.cpu().numpy(), W, S)
    assert plan.num_windows == len(ref["h"]) == {1: n, 100000: 1000, 20000: 4996}[S]
    for kg, kr in (("start_pos", "start"), ("end_pos", "end"), ("mid_pos", "mid"), ("nhet", "nhet"),
                   ("nonmissing", "nonmissing"), ("het", "h")):
        P.assert_exact(res[kg], ref[kr], kg)


def test_c4_fst_3e9_sites_properties(pgt):
    import torch
    from popgenomicstools_b200.workloads import human_like_contigs
    free, _ = torch.cuda.mem_get_info()
    n_total = 3_000_000_000 if free > 75e9 else 1_000_000_000
    W, S = 50000, 10000
    names, offs = human_like_contigs(n_total, S)
    a, b = pgt.synth_fst(4, 0, n_total)
    pos = pgt.synth_pos(4, 0, n_total, offs, 1)
    plan = pgt.WindowPlan(offs, W, S)
    res = npy(pgt.fst_window(plan, pos, a, b))
    first, last, label = plan.windows()
    # enumeration: counts, labels, positions (pos = local index + 1), uint32 midpoint
    P.assert_exact(res["nsites"], (last - first + 1).astype(np.uint32), "nsites")
    P.assert_exact(res["label"], label, "label")
    cf = np.searchsorted(offs, first, side="right") - 1
    P.assert_exact(res["start_pos"], (first - offs[cf] + 1).astype(np.uint32), "start")
    P.assert_exact(res["end_pos"], (last - offs[label] + 1).astype(np.uint32), "end")
    P.assert_exact(res["mid_pos"], ((res["start_pos"].astype(np.uint64) + res["end_pos"]) // 2).astype(np.uint32), "mid")
    carried = np.nonzero(cf != label)[0]
    assert len(carried) >= 4, "contig 1 was built to trigger the cross-contig carry"
    # sampled windows (incl. carried ones and the trailing partials) vs exactly rounded sums
    rng = np.random.default_rng(0)
    sample = np.unique(np.concatenate([rng.integers(0, plan.num_windows, 40), carried[:4],
                                       np.nonzero(res["nsites"] < W)[0][:6], [0, plan.num_windows - 1]]))
    for w in sample:
        sa = a[int(first[w]):int(last[w]) + 1].cpu().numpy()
        sb = b[int(first[w]):int(last[w]) + 1].cpu().numpy()
        ea, eb = math.fsum(sa), math.fsum(sb)
        assert abs(res["sum_a"][w] - ea) <= 1e-9 * abs(ea) + 1e-12 * np.abs(sa).sum(), w
        assert abs(res["sum_b"][w] - eb) <= 1e-9 * abs(eb) + 1e-12 * np.abs(sb).sum(), w
        assert res["fst"][w] == (res["sum_a"][w] / res["sum_b"][w] if res["sum_b"][w] != 0 else 0.0)
    # tiling identity: a 50000-site window is the union of five consecutive 10000-site windows
    plan1 = pgt.WindowPlan(offs, S, S)
    r1 = npy(pgt.fst_window(plan1, pos, a, b))
    f1, l1, _ = plan1.windows()
    idx = {int(f): i for i, f in enumerate(f1)}
    for w in sample:
        if res["nsites"][w] != W or cf[w] != label[w]:
            continue
        i0 = idx[int(first[w])]
        assert l1[i0 + 4] == last[w]
        for k in ("sum_a", "sum_b"):
            parts = r1[k][i0:i0 + 5]
            assert abs(parts.sum() - res[k][w]) <= 1e-11 * np.abs(parts).sum(), (w, k)
    # checksum of checksums: all disjoint S-windows together = the whole genome (exactly rounded per window above;
    # here against torch's own fp64 reduction of the columns)
    tot_a, tot_b = float(a.sum().item()), float(b.sum().item())
    in_windows = np.ones(len(f1), bool)
    covered_a = r1["sum_a"].sum()
    # sites of EOF-dropped partials are not in any window; add them back from the raw column
    gaps = np.concatenate([[0], l1 + 1])[:-1] != f1
    assert not gaps.any(), "disjoint windows tile the axis without holes"
    tail = a[int(l1[-1]) + 1:].sum().item() if int(l1[-1]) + 1 < n_total else 0.0
    assert abs(covered_a + tail - tot_a) <= 1e-9 * abs(tot_a) + 1e-12 * float(a.abs().sum().item())
    tail_b = b[int(l1[-1]) + 1:].sum().item() if int(l1[-1]) + 1 < n_total else 0.0
    assert abs(r1["sum_b"].sum() + tail_b - tot_b) <= 1e-9 * abs(tot_b)
    # both level-1 kernels, and 8 shards, reproduce the table bit for bit
    try:
        pgt.tune("level1", 1)
        rd = npy(pgt.fst_window(plan, pos, a, b))
    finally:
        pgt.tune("level1", 0)
    for k in res:
        assert rd[k].tobytes() == res[k].tobytes(), k
    parts = []
    for r in range(8):
        wl, wh, sl, sh = plan.shard(r, 8)
        o = pgt.fst_window(plan, pos[sl:sh], a[sl:sh], b[sl:sh], window_range=(wl, wh), site_origin=sl)
        parts.append(npy(o))
    for k in res:
        assert np.concatenate([p[k] for p in parts]).tobytes() == res[k].tobytes(), k
    # the bench / CLI configuration (512-site reduction units): same table, sums within 1e-12, and again
    # bit-identical across the level-1 kernels and 8 shards
    plan512 = pgt.WindowPlan(offs, W, S, unit_sites=512)
    r5 = npy(pgt.fst_window(plan512, pos, a, b))
    for k in ("label", "start_pos", "end_pos", "mid_pos", "nsites"):
        P.assert_exact(r5[k], res[k], k)
    for k in ("sum_a", "sum_b", "fst"):
        assert np.all(np.abs(r5[k] - res[k]) <= 1e-12 * np.abs(res[k]) + 1e-15), k
    parts = []
    for r in range(8):
        wl, wh, sl, sh = plan512.shard(r, 8)
        parts.append(npy(pgt.fst_window(plan512, pos[sl:sh], a[sl:sh], b[sl:sh], window_range=(wl, wh), site_origin=sl)))
    for k in r5:
        assert np.concatenate([p[k] for p in parts]).tobytes() == r5[k].tobytes(), k


def _exact_window_sums(cols, first, last, minind):
    """Exactly rounded (math.fsum) / exact integer statistics of the sites [first, last] from the raw columns."""
    sl = slice(int(first), int(last) + 1)
    a, b = cols["a"][sl].cpu().numpy(), cols["b"][sl].cpu().numpy()
    g = cols["geno"][sl].cpu().numpy()
    f1, f2 = cols["f1"][sl].cpu().numpy(), cols["f2"][sl].cpu().numpy()
    n1, n2 = cols["n1"][sl].cpu().numpy(), cols["n2"][sl].cpu().numpy()
    ok = (n1 >= minind) & (n2 >= minind)
    v = f1 * (1.0 - f2) + f2 * (1.0 - f1)  # numpy does not contract to FMA: dxyWindow.cpp:381 bit for bit
    return dict(sum_a=math.fsum(a), sum_b=math.fsum(b), abs_a=float(np.abs(a).sum()), abs_b=float(np.abs(b).sum()),
                nhet=int((g == 1).sum()), nonmissing=int((g >= 0).sum()), dxy=math.fsum(v[ok]), neffective=int(ok.sum()),
                nskip=int((~ok).sum()))


def _check_samples(res, cols, first, last, sample, minind):
    for w in sample:
        e = _exact_window_sums(cols, first[w], last[w], minind)
        assert abs(res["sum_a"][w] - e["sum_a"]) <= 1e-9 * abs(e["sum_a"]) + 1e-12 * e["abs_a"], w
        assert abs(res["sum_b"][w] - e["sum_b"]) <= 1e-9 * abs(e["sum_b"]) + 1e-12 * e["abs_b"], w
        assert abs(res["dxy"][w] - e["dxy"]) <= 1e-9 * abs(e["dxy"]) + 1e-300, w
        for k in ("nhet", "nonmissing", "neffective", "nskip"):
            assert int(res[k][w]) == e[k], (w, k)
        assert res["fst"][w] == (res["sum_a"][w] / res["sum_b"][w] if res["sum_b"][w] != 0 else 0.0)
        assert res["het"][w] == (res["nhet"][w] / res["nonmissing"][w] if res["nonmissing"][w] else 0.0)


def _fused_columns(pgt, seed, n, offs):
    a, b = pgt.synth_fst(seed, 0, n)
    f1, f2, n1, n2 = pgt.synth_dxy(seed, 0, n)
    return dict(pos=pgt.synth_pos(seed, 0, n, offs, 1), a=a, b=b, geno=pgt.synth_het(seed, 0, n), f1=f1, f2=f2, n1=n1, n2=n2)


_ORDER = ("pos", "a", "b", "geno", "f1", "f2", "n1", "n2")


def _integer_stats_of_every_window(torch, cols, first, last, minind):
    """nhet / nonmissing / neffective / nskip of EVERY window from prefix counts on the device.  The prefix counts
    are int32 and may wrap at 3e9 sites; a window holds <= W sites, so the wrapped difference is still exact."""
    fi = torch.from_numpy(first.astype(np.int64)).cuda()
    la = torch.from_numpy(last.astype(np.int64)).cuda()
    ok = (cols["n1"] >= minind) & (cols["n2"] >= minind)
    out = {}
    for name, flag in (("nhet", cols["geno"] == 1), ("nonmissing", cols["geno"] >= 0), ("neffective", ok), ("nskip", ~ok)):
        c = torch.cumsum(flag, 0, dtype=torch.int32)
        out[name] = (c[la] - c[fi] + flag[fi].to(torch.int32)).cpu().numpy().astype(np.int64)
        del c, flag
    return out


def test_c5_fused_3e9_sites_properties(pgt):
    """BASELINE config 5 at its named size: fused fst + dxy + het, 3e9 sites over 24 contigs, W = 1000, S = 100.
    There is no fused reference tool: the target is "each statistic equals what fstWindow.cpp:69-107,
    hetWindow.cpp:66-105 and dxyWindow.cpp:172-209,381 print" -- checked on sampled windows against exactly
    rounded sums of their sites, on EVERY window for the integer statistics, through the tiling identity, the
    equality fused == three single-statistic scans, and 8-shard invariance."""
    import torch
    from popgenomicstools_b200.workloads import human_like_contigs
    free, _ = torch.cuda.mem_get_info()
    n_total = 3_000_000_000 if free > 160e9 else (1_000_000_000 if free > 60e9 else 300_000_000)
    W, S, minind = 1000, 100, 5
    names, offs = human_like_contigs(n_total, S)
    cols = _fused_columns(pgt, 5, n_total, offs)
    plan = pgt.WindowPlan(offs, W, S)
    res = npy(pgt.fused_window(plan, *[cols[k] for k in _ORDER], minind=minind))
    first, last, label = plan.windows()
    assert plan.num_windows == len(first) and abs(plan.num_windows - n_total // S) < 400
    P.assert_exact(res["nsites"], (last - first + 1).astype(np.uint32), "nsites")
    P.assert_exact(res["label"], label, "label")
    cf = np.searchsorted(offs, first, side="right") - 1
    P.assert_exact(res["start_pos"], (first - offs[cf] + 1).astype(np.uint32), "start")
    P.assert_exact(res["end_pos"], (last - offs[label] + 1).astype(np.uint32), "end")
    P.assert_exact(res["mid_pos"], ((res["start_pos"].astype(np.uint64) + res["end_pos"]) // 2).astype(np.uint32), "mid")
    carried = np.nonzero(cf != label)[0]
    assert len(carried) >= 9, "contig 1 was built to trigger the cross-contig carry"
    rng = np.random.default_rng(5)
    sample = np.unique(np.concatenate([rng.integers(0, plan.num_windows, 60), carried[:4], np.nonzero(res["nsites"] < W)[0][:6],
                                       [0, plan.num_windows - 1]]))
    _check_samples(res, cols, first, last, sample, minind)
    every = _integer_stats_of_every_window(torch, cols, first, last, minind)
    for k, v in every.items():
        P.assert_exact(res[k].astype(np.int64), v, k + " (every window)")
    # fused == the three single-statistic scans, bit for bit
    for one in (pgt.fst_window(plan, cols["pos"], cols["a"], cols["b"]), pgt.het_window(plan, cols["pos"], cols["geno"]),
                pgt.dxy_window(plan, cols["pos"], cols["f1"], cols["f2"], cols["n1"], cols["n2"], minind=minind)):
        o = npy(one)
        for k, v in o.items():
            assert v.tobytes() == res[k].tobytes(), k
        del one, o
    # tiling identity: a full 1000-site window is the union of ten disjoint 100-site windows
    plan1 = pgt.WindowPlan(offs, S, S)
    r1 = npy(pgt.fst_window(plan1, cols["pos"], cols["a"], cols["b"]))
    f1w, l1w, _ = plan1.windows()
    for w in sample:
        if res["nsites"][w] != W or cf[w] != label[w]:
            continue
        i0 = int(np.searchsorted(f1w, first[w]))
        assert f1w[i0] == first[w] and l1w[i0 + 9] == last[w]
        for k in ("sum_a", "sum_b"):
            parts = r1[k][i0:i0 + 10]
            assert abs(parts.sum() - res[k][w]) <= 1e-11 * np.abs(parts).sum(), (w, k)
    del r1
    # 8 shards reproduce the table bit for bit; the shards' global lines add up to the unsharded one
    parts = []
    for r in range(8):
        wl, wh, sl, sh = plan.shard(r, 8)
        parts.append(npy(pgt.fused_window(plan, *[cols[k][sl:sh] for k in _ORDER], minind=minind, window_range=(wl, wh), site_origin=sl)))
    for k in res:
        if k != "dxy_global":
            assert np.concatenate([p[k] for p in parts]).tobytes() == res[k].tobytes(), k
    gl = np.sum([p["dxy_global"] for p in parts], axis=0)
    assert gl[1] == res["dxy_global"][1] and gl[2] == res["dxy_global"][2]
    assert abs(gl[0] - res["dxy_global"][0]) <= 1e-12 * res["dxy_global"][0]
    assert res["dxy_global"][1] + res["dxy_global"][2] == n_total


def test_c5_stress_variant_s1_1e8_sites(pgt):
    """SURVEY 8(d) C5 stress variant: W = 1000, S = 1 on a 1e8-site subset (one window per site, output as large
    as the input).  Sliding-tile path.  Oracle on a 1e6-site prefix (every window of it), the integer statistics of
    EVERY window exactly, sampled windows against exactly rounded sums, fused == single scans, 8 shards."""
    import torch
    from popgenomicstools_b200 import _cabi
    from popgenomicstools_b200.workloads import human_like_contigs
    n_total, W, S, minind = 100_000_000, 1000, 1, 5
    names, offs = human_like_contigs(n_total, S)
    cols = _fused_columns(pgt, 5, n_total, offs)
    plan = pgt.WindowPlan(offs, W, S)
    assert plan.scan_path(_cabi.PGT_STAT_FUSED) == "slide"
    assert plan.num_windows == n_total - W + 1  # S = 1: the buffer is full at every contig change, one segment
    res = npy(pgt.fused_window(plan, *[cols[k] for k in _ORDER], minind=minind))
    first, last, label = plan.windows()
    P.assert_exact(res["nsites"], (last - first + 1).astype(np.uint32), "nsites")
    P.assert_exact(res["label"], label, "label")
    cf = np.searchsorted(offs, first, side="right") - 1
    P.assert_exact(res["start_pos"], (first - offs[cf] + 1).astype(np.uint32), "start")
    P.assert_exact(res["end_pos"], (last - offs[label] + 1).astype(np.uint32), "end")
    P.assert_exact(res["mid_pos"], ((res["start_pos"].astype(np.uint64) + res["end_pos"]) // 2).astype(np.uint32), "mid")
    # oracle (the reference's re-sum + slide loop restated) on a prefix: its full windows are the first windows here
    npre = 1_000_000
    assert int(offs[1]) > npre
    h = {k: v[:npre].cpu().numpy() for k, v in cols.items()}
    chr0 = np.zeros(npre, np.uint32)
    rf = O.fst(chr0, h["pos"], h["a"], h["b"], W, S)
    ra = O.fst(chr0, h["pos"], np.abs(h["a"]), np.abs(h["b"]), W, S)
    rh = O.het(chr0, h["pos"], h["geno"], W, S)
    rd = O.dxy(chr0, h["pos"], h["f1"], h["f2"], h["n1"], h["n2"], minind, W, S, 1)
    m = npre - W + 1
    assert len(rf["n"]) >= m
    P.assert_exact(res["start_pos"][:m], rf["start"][:m], "start (prefix)")
    P.assert_exact(res["end_pos"][:m], rf["end"][:m], "end (prefix)")
    P.assert_sum_close(res["sum_a"][:m], rf["asum"][:m], ra["asum"][:m], "sum_a (prefix)")
    P.assert_sum_close(res["sum_b"][:m], rf["bsum"][:m], ra["bsum"][:m], "sum_b (prefix)")
    P.assert_exact(res["nhet"][:m], rh["nhet"][:m], "nhet (prefix)")
    P.assert_exact(res["nonmissing"][:m], rh["nonmissing"][:m], "nonmissing (prefix)")
    P.assert_exact(res["het"][:m], rh["h"][:m], "het (prefix)")
    P.assert_exact(res["neffective"][:m], rd["neff"][:m], "neff (prefix)")
    P.assert_exact(res["nskip"][:m], rd["nskip"][:m], "nskip (prefix)")
    P.assert_sum_close(res["dxy"][:m], rd["dxy"][:m], rd["dxy"][:m], "dxy (prefix)")
    # every window: integer statistics; samples: exactly rounded sums (incl. windows straddling contigs)
    every = _integer_stats_of_every_window(torch, cols, first, last, minind)
    for k, v in every.items():
        P.assert_exact(res[k].astype(np.int64), v, k + " (every window)")
    del every
    rng = np.random.default_rng(6)
    straddle = np.nonzero(cf != label)[0]
    sample = np.unique(np.concatenate([rng.integers(0, plan.num_windows, 60), straddle[:3], straddle[-3:], [0, plan.num_windows - 1]]))
    _check_samples(res, cols, first, last, sample, minind)
    del first, last, label, cf
    # fused == single-statistic scans bit for bit
    o = npy(pgt.fst_window(plan, cols["pos"], cols["a"], cols["b"]))
    for k, v in o.items():
        assert v.tobytes() == res[k].tobytes(), k
    del o
    o = npy(pgt.het_window(plan, cols["pos"], cols["geno"]))
    for k, v in o.items():
        assert v.tobytes() == res[k].tobytes(), k
    del o
    # 8 shards, bit for bit (fst columns: the table is 7.6 GB when fused)
    ref = {k: res[k] for k in ("sum_a", "sum_b", "fst", "start_pos", "end_pos", "label", "nsites")}
    del res
    at = 0
    for r in range(8):
        wl, wh, sl, sh = plan.shard(r, 8)
        p = npy(pgt.fst_window(plan, cols["pos"][sl:sh], cols["a"][sl:sh], cols["b"][sl:sh], window_range=(wl, wh), site_origin=sl))
        assert wl == at
        for k, v in ref.items():
            assert p[k].tobytes() == v[wl:wh].tobytes(), (r, k)
        at = wh
    assert at == plan.num_windows
