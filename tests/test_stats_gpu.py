"""hetWindow / dxyWindow (-fixedsite 1 and bp mode) / fused hot paths on the GPU, through the
C ABI, vs the oracle and the committed reference transcripts; device- and host-memory modes."""
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


def to_dev(x, dtype):
    import torch
    x = np.ascontiguousarray(x, dtype)
    if dtype == np.uint32:
        return torch.from_numpy(x.view(np.int32)).cuda().view(torch.uint32)
    return torch.from_numpy(x).cuda()


def npy(out):
    import torch
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}


def offsets(lengths):
    return np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)


# ------------------------------------------------------------------------------------ het

def check_het(res, lengths, pos, g, W, S):
    ref = O.het(T.expand_chr(lengths), pos, g, W, S)
    for kg, kr in (("label", "label"), ("start_pos", "start"), ("end_pos", "end"), ("mid_pos", "mid"),
                   ("nhet", "nhet"), ("nonmissing", "nonmissing")):
        P.assert_exact(res[kg], ref[kr], kg)
    P.assert_exact(res["het"], ref["h"], "het (integer counts then one IEEE divide: bit-exact)")
    P.assert_exact(res["nsites"], ref["last"] - ref["first"] + 1, "nsites")
    return ref


def test_het_golden(pgt, golden_cases):
    n = 0
    for c in golden_cases:
        if c["tool"] != "hetWindow":
            continue
        g = np.clip(np.asarray(c["geno"]), -128, 127).astype(np.int8)
        pos = np.asarray(c["pos"], np.uint32)
        plan = pgt.WindowPlan(offsets(c["lengths"]), c["W"], c["S"], unit_sites=32)
        res = npy(pgt.het_window(plan, to_dev(pos, np.uint32), to_dev(g, np.int8)))
        check_het(res, c["lengths"], pos, g, c["W"], c["S"])
        rows = O.het_rows(dict(label=res["label"], start=res["start_pos"], end=res["end_pos"], mid=res["mid_pos"],
                               h=res["het"], nonmissing=res["nonmissing"]), c["names"])
        assert rows == c["stdout"].splitlines()
        # host-memory mode gives the same answer
        resh = pgt.het_window(plan, pos, g)
        for k in res:
            assert np.array_equal(resh[k], res[k]), k
        n += 1
    assert n >= 30


@pytest.mark.parametrize("W,S", [(1, 1), (100000, 100000), (100000, 20000), (1000, 100), (37, 5)])
def test_het_synthetic(pgt, W, S):
    """BASELINE config 3 shapes (single-site and 100 kb windows) on a scaled chromosome + carry contigs."""
    n_total = 1_000_000 if W > 1 else 300_000
    lengths = [W + 2 * S if W + 2 * S < n_total // 2 else n_total // 3]
    lengths += [n_total - lengths[0] - 11, 11]
    offs = offsets(lengths)
    n = int(offs[-1])
    g = O.synth_het(3, 0, n)
    pos = O.synth_pos(3, offs, 1)
    gg = pgt.synth_het(3, 0, n)
    assert np.array_equal(gg.cpu().numpy(), g)
    plan = pgt.WindowPlan(offs, W, S)
    res = npy(pgt.het_window(plan, to_dev(pos, np.uint32), gg))
    check_het(res, lengths, pos, g, W, S)


# ------------------------------------------------------------------------------------ dxy

def check_dxy(res, ref, absref, skip_missing=0):
    if skip_missing:
        keep = res["neffective"] > 0
        res = {k: (v[keep] if k != "dxy_global" else v) for k, v in res.items()}
    P.assert_exact(res["label"], ref["label"], "label")
    P.assert_exact(res["start_pos"].astype(np.int64), ref["start"], "start")
    P.assert_exact(res["end_pos"].astype(np.int64), ref["end"], "end")
    P.assert_exact(res["neffective"], ref["neff"], "neffective")
    P.assert_exact(res["nskip"], ref["nskip"], "nskip")
    P.assert_sum_close(res["dxy"], ref["dxy"], absref["dxy"], "dxy")
    g, rg = res["dxy_global"], ref["global"]
    assert g[1] == rg[1] and g[2] == rg[2], (g, rg)
    assert abs(g[0] - rg[0]) <= 1e-9 * abs(rg[0]) + 1e-12 * abs(rg[0]), (g, rg)
    return res


def run_dxy(pgt, lengths, pos, f1, f2, n1, n2, minind, W, S, fixedsite, chr_len, host=False, unit=0):
    soff = offsets(lengths)
    if fixedsite:
        plan = pgt.WindowPlan(soff, W, S, unit_sites=unit)
        so = None
    else:
        plan = pgt.WindowPlan(offsets(chr_len), W, S, mode="bp", unit_sites=unit)
        so = soff
    cols = [np.ascontiguousarray(pos, np.uint32), np.ascontiguousarray(f1, np.float64), np.ascontiguousarray(f2, np.float64),
            np.ascontiguousarray(n1, np.int32), np.ascontiguousarray(n2, np.int32)]
    if not host:
        cols = [to_dev(c, c.dtype.type) for c in cols]
    return plan, npy(pgt.dxy_window(plan, *cols, minind=minind, site_offsets=so))


def test_dxy_golden(pgt, golden_cases):
    n = ties = 0
    for c in golden_cases:
        if c["tool"] != "dxyWindow":
            continue
        f1, f2 = T.micro_to_f64(c["f1_micro"]), T.micro_to_f64(c["f2_micro"])
        pos = np.asarray(c["pos"], np.uint32)
        chr_id = T.expand_chr(c["lengths"])
        W, S = c["W"], c["S"]
        ref = O.dxy(chr_id, pos, f1, f2, c["n1"], c["n2"], c["minind"], W, S, c["fixedsite"], c["skip_missing"], c["chr_len"])
        if W == 0:
            # global mode (-winsize 0): only the global line; any site-mode plan yields it
            plan, res = run_dxy(pgt, c["lengths"], pos, f1, f2, c["n1"], c["n2"], c["minind"], 1, 1, 1, None)
            g = res["dxy_global"]
            row = "\t".join([O.g6(g[0]), str(int(g[1])), str(int(g[2]))])
            assert P.rows_match_modulo_ties([row], c["stdout"].splitlines(), {0}) <= 1
            n += 1
            continue
        absref = O.dxy(chr_id, pos, f1, f2, c["n1"], c["n2"], c["minind"], W, S, c["fixedsite"], c["skip_missing"], c["chr_len"])
        for host in (False, True):
            plan, res = run_dxy(pgt, c["lengths"], pos, f1, f2, c["n1"], c["n2"], c["minind"], W, S, c["fixedsite"],
                                c["chr_len"], host=host, unit=32)
            res = check_dxy(res, ref, absref, c["skip_missing"])
        rows = O.dxy_rows(dict(label=res["label"], start=res["start_pos"], end=res["end_pos"], dxy=res["dxy"],
                               neff=res["neffective"], nskip=res["nskip"]), c["names"])
        ties += P.rows_match_modulo_ties(rows, c["stdout"].splitlines(), {3})
        n += 1
    assert n >= 80
    assert ties <= 3, ties


@pytest.mark.parametrize("fixedsite,W,S,density", [(0, 20000, 5000, 10), (0, 20000, 5000, 1), (1, 20000, 5000, 10),
                                                   (0, 1000, 1000, 3), (1, 1, 1, 10), (0, 777, 100, 10)])
def test_dxy_synthetic(pgt, fixedsite, W, S, density):
    """BASELINE config 2 shape (20 kb windows / 5 kb step over two MAFs) on a scaled contig set:
    sparse (1 site per 10 bp) and dense (every bp), bp and fixed-site windows."""
    nsites = [60_000, 35_000, 900, 17_000]
    soff = offsets(nsites)
    n = int(soff[-1])
    chr_len = [L * density + (37 if i % 2 else 0) for i, L in enumerate(nsites)]
    chr_len[0] = max(W, (chr_len[0] // S) * S)  # (L-W)%S==0 -> bp-axis carry into chromosome 2
    if density > 1:
        nsites[0] = min(nsites[0], chr_len[0] // density)
        soff = offsets(nsites)
        n = int(soff[-1])
    f1, f2, n1, n2 = O.synth_dxy(2, 0, n)
    pos = O.synth_pos(2, soff, density)
    for c in range(len(nsites)):
        assert pos[soff[c + 1] - 1] <= chr_len[c]
    gf1, gf2, gn1, gn2 = pgt.synth_dxy(2, 0, n)
    gpos = pgt.synth_pos(2, 0, n, soff, density)
    assert np.array_equal(gf1.cpu().numpy(), f1) and np.array_equal(gf2.cpu().numpy(), f2)
    assert np.array_equal(gn1.cpu().numpy(), n1) and np.array_equal(gn2.cpu().numpy(), n2)
    assert np.array_equal(gpos.cpu().numpy(), pos)
    chr_id = T.expand_chr(nsites)
    ref = O.dxy(chr_id, pos, f1, f2, n1, n2, 5, W, S, fixedsite, 0, chr_len)
    absref = ref  # per-site dxy >= 0: sum|x| == sum x
    for host in (False, True):
        plan, res = run_dxy(pgt, nsites, pos, f1, f2, n1, n2, 5, W, S, fixedsite, chr_len, host=host)
        check_dxy(res, ref, absref)


# ------------------------------------------------------------------------------------ fused

@pytest.mark.parametrize("W,S", [(1000, 100), (1000, 1), (50000, 10000)])
def test_fused_equals_three_tools(pgt, W, S):
    """BASELINE config 5: each statistic of the fused scan equals what the respective reference
    tool computes for the same sites / W / S (fused vs oracle, and fused vs the single-stat scans
    bit for bit)."""
    n_total = 300_000 if S > 1 else 40_000
    lengths = [W + 3 * S, n_total - W - 3 * S - 13, 13]
    offs = offsets(lengths)
    n = int(offs[-1])
    chr_id = T.expand_chr(lengths)
    a, b = pgt.synth_fst(5, 0, n)
    g = pgt.synth_het(5, 0, n)
    f1, f2, n1, n2 = pgt.synth_dxy(5, 0, n)
    pos = pgt.synth_pos(5, 0, n, offs, 1)
    plan = pgt.WindowPlan(offs, W, S)
    res = npy(pgt.fused_window(plan, pos, a, b, g, f1, f2, n1, n2, minind=5))
    one_f = npy(pgt.fst_window(plan, pos, a, b))
    one_h = npy(pgt.het_window(plan, pos, g))
    one_d = npy(pgt.dxy_window(plan, pos, f1, f2, n1, n2, minind=5))
    for one in (one_f, one_h, one_d):
        for k, v in one.items():
            assert v.tobytes() == res[k].tobytes(), k
    hpos = pos.cpu().numpy()
    ha, hb = a.cpu().numpy(), b.cpu().numpy()
    ref = O.fst(chr_id, hpos, ha, hb, W, S)
    absr = O.fst(chr_id, hpos, np.abs(ha), np.abs(hb), W, S)
    P.assert_sum_close(res["sum_a"], ref["asum"], absr["asum"], "sum_a")
    P.assert_sum_close(res["sum_b"], ref["bsum"], absr["bsum"], "sum_b")
    P.assert_exact(res["start_pos"], ref["start"], "start")
    P.assert_exact(res["label"], ref["label"], "label")
    check_het(res, lengths, hpos, g.cpu().numpy(), W, S)
    dref = O.dxy(chr_id, hpos, f1.cpu().numpy(), f2.cpu().numpy(), n1.cpu().numpy(), n2.cpu().numpy(), 5, W, S, 1)
    P.assert_exact(res["neffective"], dref["neff"], "neff")
    P.assert_exact(res["nskip"], dref["nskip"], "nskip")
    P.assert_sum_close(res["dxy"], dref["dxy"], dref["dxy"], "dxy")


def test_fst_host_mode_matches_device_mode(pgt):
    lengths = [241170, 130000, 1234]
    offs = offsets(lengths)
    n, W, S = int(offs[-1]), 5000, 1000
    a, b = O.synth_fst(4, 0, n)
    pos = O.synth_pos(4, offs, 1)
    plan = pgt.WindowPlan(offs, W, S)
    dev = npy(pgt.fst_window(plan, to_dev(pos, np.uint32), to_dev(a, np.float64), to_dev(b, np.float64)))
    host = pgt.fst_window(plan, pos, a, b)
    for k in dev:
        assert host[k].tobytes() == dev[k].tobytes(), k


@pytest.mark.parametrize("W,S", [(1000, 1), (777, 13), (64, 1), (5000, 3), (40, 40)])
def test_level2_scan_mode_matches_oracle_and_direct_mode(pgt, W, S):
    """Fine steps with long windows run level 2 in scan mode (block prefix/suffix scans,
    window = SUF[first] + PRE[last]); forced on and off it must agree with the oracle, exactly
    for the integer statistics and within the stated tolerance for the sums."""
    lengths = [W + 5 * S, 20011, 3, W // 2 + 1, 9000]
    offs = offsets(lengths)
    n = int(offs[-1])
    chr_id = T.expand_chr(lengths)
    a, b = pgt.synth_fst(6, 0, n)
    g = pgt.synth_het(6, 0, n)
    f1, f2, n1, n2 = pgt.synth_dxy(6, 0, n)
    pos = pgt.synth_pos(6, 0, n, offs, 1)
    hpos, ha, hb = pos.cpu().numpy(), a.cpu().numpy(), b.cpu().numpy()
    ref = O.fst(chr_id, hpos, ha, hb, W, S)
    absr = O.fst(chr_id, hpos, np.abs(ha), np.abs(hb), W, S)
    href = O.het(chr_id, hpos, g.cpu().numpy(), W, S)
    dref = O.dxy(chr_id, hpos, f1.cpu().numpy(), f2.cpu().numpy(), n1.cpu().numpy(), n2.cpu().numpy(), 5, W, S, 1)
    plan = pgt.WindowPlan(offs, W, S)
    try:
        for mode in (2, 1, 0):
            pgt.tune("level2", mode)
            res = npy(pgt.fused_window(plan, pos, a, b, g, f1, f2, n1, n2, minind=5))
            P.assert_exact(res["label"], ref["label"], "label")
            P.assert_exact(res["start_pos"], ref["start"], "start")
            P.assert_exact(res["end_pos"], ref["end"], "end")
            P.assert_exact(res["nsites"], ref["n"], "nsites")
            P.assert_sum_close(res["sum_a"], ref["asum"], absr["asum"], f"sum_a mode {mode}")
            P.assert_sum_close(res["sum_b"], ref["bsum"], absr["bsum"], f"sum_b mode {mode}")
            P.assert_exact(res["nhet"], href["nhet"], "nhet")
            P.assert_exact(res["nonmissing"], href["nonmissing"], "nonmissing")
            P.assert_exact(res["het"], href["h"], "het")
            P.assert_exact(res["neffective"], dref["neff"], "neff")
            P.assert_exact(res["nskip"], dref["nskip"], "nskip")
            P.assert_sum_close(res["dxy"], dref["dxy"], dref["dxy"], f"dxy mode {mode}")
            assert res["dxy_global"][1] == dref["global"][1] and res["dxy_global"][2] == dref["global"][2]
            # shards in scan mode reproduce the unsharded scan-mode table bit for bit
            if mode == 2:
                parts = []
                for r in range(3):
                    wl, wh, sl, sh = plan.shard(r, 3)
                    if wh > wl:
                        parts.append(npy(pgt.fst_window(plan, pos[sl:sh], a[sl:sh], b[sl:sh], window_range=(wl, wh), site_origin=sl)))
                for k in ("sum_a", "sum_b", "fst", "start_pos", "nsites"):
                    assert np.concatenate([p[k] for p in parts]).tobytes() == res[k].tobytes(), k
    finally:
        pgt.tune("level2", 0)
