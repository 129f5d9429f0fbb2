"""Every kernel path x statistic x memory mode on one small genome, each compared with the oracle:
level 1 direct / tiled (TMA-staged) / vectorised het, level 2 warp / thread-per-window / scan mode /
per-site, unaligned column views, shards, host staging, bp mode.  (compute-sanitizer is closed on
this pool, so memory-safety evidence is this matrix: any out-of-range read of a tile, head/tail
byte or halo would change a compared value.)"""
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
    yield m
    m.tune("level1", 0)
    m.tune("level2", 0)


def offsets(l):
    return np.concatenate([[0], np.cumsum(l)]).astype(np.uint64)


def npy(out):
    import torch
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}


@pytest.mark.parametrize("W,S,u", [(5000, 100, 0), (1000, 100, 0), (300, 299, 0), (1, 1, 0), (64, 1, 0), (777, 13, 32),
                                   (2560, 256, 0), (4096, 4096, 4096), (100, 10, 0), (24, 3, 32), (8, 4, 0), (31, 1, 0), (5000, 1000, 512)])
def test_site_mode_all_paths(pgt, W, S, u):
    import torch
    lengths = [W + 3 * S, 12345, 7, 4001]
    offs = offsets(lengths)
    n = int(offs[-1])
    chr_id = T.expand_chr(lengths)
    a, b = pgt.synth_fst(1, 0, n)
    g = pgt.synth_het(1, 0, n)
    f1, f2, n1, n2 = pgt.synth_dxy(1, 0, n)
    pos = pgt.synth_pos(1, 0, n, offs, 1)
    # column views that start at odd element offsets (head/tail handling of the bulk copies / vector loads)
    A = torch.empty(n + 3, dtype=torch.float64, device="cuda")
    A[3:].copy_(a)
    G = torch.empty(n + 5, dtype=torch.int8, device="cuda")
    G[5:].copy_(g)
    h = {k: v.cpu().numpy() for k, v in dict(pos=pos, a=a, b=b, g=g, f1=f1, f2=f2, n1=n1, n2=n2).items()}
    rf = O.fst(chr_id, h["pos"], h["a"], h["b"], W, S)
    ra = O.fst(chr_id, h["pos"], np.abs(h["a"]), np.abs(h["b"]), W, S)
    rh = O.het(chr_id, h["pos"], h["g"], W, S)
    rd = O.dxy(chr_id, h["pos"], h["f1"], h["f2"], h["n1"], h["n2"], 5, W, S, 1)

    def check(res, what):
        P.assert_exact(res["label"], rf["label"], what + " label")
        P.assert_exact(res["start_pos"], rf["start"], what + " start")
        P.assert_exact(res["end_pos"], rf["end"], what + " end")
        P.assert_exact(res["nsites"], rf["n"], what + " nsites")
        if "sum_a" in res:
            P.assert_sum_close(res["sum_a"], rf["asum"], ra["asum"], what + " sum_a")
            P.assert_sum_close(res["sum_b"], rf["bsum"], ra["bsum"], what + " sum_b")
        if "nhet" in res:
            P.assert_exact(res["nhet"], rh["nhet"], what + " nhet")
            P.assert_exact(res["nonmissing"], rh["nonmissing"], what + " nonmissing")
            P.assert_exact(res["het"], rh["h"], what + " het")
        if "dxy" in res:
            P.assert_exact(res["neffective"], rd["neff"], what + " neff")
            P.assert_exact(res["nskip"], rd["nskip"], what + " nskip")
            P.assert_sum_close(res["dxy"], rd["dxy"], rd["dxy"], what + " dxy")
            assert res["dxy_global"][1] == rd["global"][1] and res["dxy_global"][2] == rd["global"][2], what

    plan = pgt.WindowPlan(offs, W, S, unit_sites=u)
    same = {}
    for l1 in (0, 1, 2):
        for l2 in (0, 1, 2):
            pgt.tune("level1", l1)
            pgt.tune("level2", l2)
            tag = f"l1={l1} l2={l2}"
            res = dict(fst=npy(pgt.fst_window(plan, pos, A[3:], b)), het=npy(pgt.het_window(plan, pos, G[5:])),
                       dxy=npy(pgt.dxy_window(plan, pos, f1, f2, n1, n2, minind=5)),
                       fused=npy(pgt.fused_window(plan, pos, A[3:], b, G[5:], f1, f2, n1, n2, minind=5)))
            for name, r in res.items():
                check(r, f"{name} {tag}")
                same[(name, l1, l2)] = r
    keys = lambda r: [k for k in r if k != "dxy_global"]
    for name in ("fst", "het", "dxy", "fused"):
        base = same[(name, 0, 1)]       # warp per window
        scan = same[(name, 0, 2)]       # scan mode (its own summation order)
        for l1 in (0, 1, 2):
            # the level-1 kernels implement one summation order: bit-identical partials, hence windows
            for l2 in (1, 2):
                r, want = same[(name, l1, l2)], (base if l2 == 1 else scan)
                assert all(r[k].tobytes() == want[k].tobytes() for k in keys(r)), f"{name} l1={l1} l2={l2} differs from l1=0"
            # auto picks one of the two level-2 orders
            r = same[(name, l1, 0)]
            assert all(r[k].tobytes() == base[k].tobytes() for k in keys(r)) or \
                all(r[k].tobytes() == scan[k].tobytes() for k in keys(r)), f"{name} l1={l1} auto matches neither order"
    pgt.tune("level1", 0)
    pgt.tune("level2", 0)
    full = npy(pgt.fused_window(plan, pos, a, b, g, f1, f2, n1, n2, minind=5))
    parts = []
    for r in range(3):
        wl, wh, sl, sh = plan.shard(r, 3)
        if wh > wl:
            parts.append(npy(pgt.fused_window(plan, pos[sl:sh], a[sl:sh], b[sl:sh], g[sl:sh], f1[sl:sh], f2[sl:sh], n1[sl:sh],
                                              n2[sl:sh], minind=5, window_range=(wl, wh), site_origin=sl)))
    for k in full:
        if k == "dxy_global":
            tot = sum(p[k] for p in parts)
            assert tot[1] == full[k][1] and tot[2] == full[k][2] and abs(tot[0] - full[k][0]) <= 1e-9 * abs(full[k][0])
        else:
            assert np.concatenate([p[k] for p in parts]).tobytes() == full[k].tobytes(), k
    hres = pgt.fused_window(plan, h["pos"], h["a"], h["b"], h["g"], h["f1"], h["f2"], h["n1"], h["n2"], minind=5)
    check(hres, "host mode")
    for k in full:
        if k != "dxy_global":
            assert hres[k].tobytes() == full[k].tobytes(), "host vs device " + k


def test_warp_per_window_rows_staged_over_many_batches(pgt):
    """k_windows stores its rows 32 at a time from a double-buffered shared-memory stage.  Here every block walks
    several batches (1e5 windows of 80 units over 1184 blocks: 32 + 32 + a ragged rest), contig ends fall inside
    batches, and the device table is compared with the oracle row by row; a shard's table (its first row at an
    odd offset of the block runs) must hold the same bits."""
    lengths = [1_700_003, 2560 + 31, 900_000, 5, 700_017]
    offs = offsets(lengths)
    n = int(offs[-1])
    W, S = 2560, 32
    chr_id = T.expand_chr(lengths)
    a, b = pgt.synth_fst(7, 0, n)
    g = pgt.synth_het(7, 0, n)
    f1, f2, n1, n2 = pgt.synth_dxy(7, 0, n)
    pos = pgt.synth_pos(7, 0, n, offs, 1)
    h = {k: v.cpu().numpy() for k, v in dict(pos=pos, a=a, b=b, g=g, f1=f1, f2=f2, n1=n1, n2=n2).items()}
    rf = O.fst(chr_id, h["pos"], h["a"], h["b"], W, S)
    ra = O.fst(chr_id, h["pos"], np.abs(h["a"]), np.abs(h["b"]), W, S)
    rh = O.het(chr_id, h["pos"], h["g"], W, S)
    rd = O.dxy(chr_id, h["pos"], h["f1"], h["f2"], h["n1"], h["n2"], 5, W, S, 1)
    plan = pgt.WindowPlan(offs, W, S, unit_sites=32)
    assert plan.num_windows > 1184 * 64 and plan.scan_path(pgt._cabi.PGT_STAT_FUSED) == "units"
    pgt.tune("level2", 1)  # warp per window
    try:
        res = npy(pgt.fused_window(plan, pos, a, b, g, f1, f2, n1, n2, minind=5))
        wl, wh, sl, sh = plan.shard(1, 3)
        part = npy(pgt.fused_window(plan, pos[sl:sh], a[sl:sh], b[sl:sh], g[sl:sh], f1[sl:sh], f2[sl:sh], n1[sl:sh], n2[sl:sh],
                                    minind=5, window_range=(wl, wh), site_origin=sl))
    finally:
        pgt.tune("level2", 0)
    assert len(res["label"]) == len(rf["label"]) == plan.num_windows
    for k, want in (("label", rf["label"]), ("start_pos", rf["start"]), ("end_pos", rf["end"]), ("mid_pos", rf["mid"]), ("nsites", rf["n"]),
                    ("nhet", rh["nhet"]), ("nonmissing", rh["nonmissing"]), ("het", rh["h"]), ("neffective", rd["neff"]),
                    ("nskip", rd["nskip"])):
        P.assert_exact(res[k], want, k)
    P.assert_sum_close(res["sum_a"], rf["asum"], ra["asum"], "sum_a")
    P.assert_sum_close(res["sum_b"], rf["bsum"], ra["bsum"], "sum_b")
    P.assert_sum_close(res["dxy"], rd["dxy"], rd["dxy"], "dxy")
    for k in res:
        if k != "dxy_global":
            assert part[k].tobytes() == res[k][wl:wh].tobytes(), k


@pytest.mark.parametrize("density", [10, 1])
@pytest.mark.parametrize("W,S", [(2000, 500), (100, 100), (1, 1), (777, 10)])
def test_bp_mode_all_paths(pgt, density, W, S):
    nsites = [3000, 1500, 40]
    soff = offsets(nsites)
    ns = int(soff[-1])
    chr_len = [x * density + 17 for x in nsites]
    f1, f2, n1, n2 = pgt.synth_dxy(2, 0, ns)
    p = pgt.synth_pos(2, 0, ns, soff, density)
    h = [x.cpu().numpy() for x in (p, f1, f2, n1, n2)]
    ref = O.dxy(T.expand_chr(nsites), h[0], h[1], h[2], h[3], h[4], 5, W, S, 0, 0, chr_len)
    plan = pgt.WindowPlan(offsets(chr_len), W, S, mode="bp")

    def check(res, what):
        P.assert_exact(res["label"], ref["label"], what)
        P.assert_exact(res["start_pos"].astype(np.int64), ref["start"], what)
        P.assert_exact(res["end_pos"].astype(np.int64), ref["end"], what)
        P.assert_exact(res["neffective"], ref["neff"], what)
        P.assert_exact(res["nskip"], ref["nskip"], what)
        P.assert_sum_close(res["dxy"], ref["dxy"], ref["dxy"], what)

    for l1 in (0, 2):
        for l2 in (0, 1, 2):
            pgt.tune("level1", l1)
            pgt.tune("level2", l2)
            check(npy(pgt.dxy_window(plan, p, f1, f2, n1, n2, minind=5, site_offsets=soff)), f"bp l1={l1} l2={l2}")
    pgt.tune("level1", 0)
    pgt.tune("level2", 0)
    check(pgt.dxy_window(plan, *h, minind=5, site_offsets=soff), "bp host")


@pytest.mark.parametrize("W,S", [(1000, 1000), (5000, 1000), (100, 10), (3000, 3000)])
def test_many_contigs_tile_segment_table(pgt, W, S):
    """Scaffold-level genomes: hundreds of contigs put consecutive tiles of a CTA into different
    segments; above 32 segments the tiled kernel uses a tile -> segment table (k_tile_segs).  Same
    bits as the direct kernel (which binary-searches per unit), oracle parity, shards, host mode."""
    import torch
    rng = np.random.default_rng(W + S)
    lengths = rng.integers(1, 9000, size=400).tolist() + [1, 2, W, W + S, 3 * W]
    offs = offsets(lengths)
    n = int(offs[-1])
    a, b = pgt.synth_fst(8, 0, n)
    pos = pgt.synth_pos(8, 0, n, offs, 1)
    plan = pgt.WindowPlan(offs, W, S)
    assert plan.num_segments > 32
    ref = O.fst(T.expand_chr(lengths), pos.cpu().numpy(), a.cpu().numpy(), b.cpu().numpy(), W, S)
    absr = O.fst(T.expand_chr(lengths), pos.cpu().numpy(), np.abs(a.cpu().numpy()), np.abs(b.cpu().numpy()), W, S)
    try:
        pgt.tune("level1", 2)  # tiled (the input is far below the automatic 256 MB threshold)
        tiled = npy(pgt.fst_window(plan, pos, a, b))
        parts = []
        for r in range(3):
            wl, wh, sl, sh = plan.shard(r, 3)
            if wh > wl:
                parts.append(npy(pgt.fst_window(plan, pos[sl:sh], a[sl:sh], b[sl:sh], window_range=(wl, wh), site_origin=sl)))
        host = pgt.fst_window(plan, pos.cpu().numpy(), a.cpu().numpy(), b.cpu().numpy())
        pgt.tune("level1", 1)
        direct = npy(pgt.fst_window(plan, pos, a, b))
    finally:
        pgt.tune("level1", 0)
    P.assert_exact(tiled["label"], ref["label"], "label")
    P.assert_exact(tiled["start_pos"], ref["start"], "start")
    P.assert_exact(tiled["end_pos"], ref["end"], "end")
    P.assert_exact(tiled["nsites"], ref["n"], "nsites")
    P.assert_sum_close(tiled["sum_a"], ref["asum"], absr["asum"], "sum_a")
    P.assert_sum_close(tiled["sum_b"], ref["bsum"], absr["bsum"], "sum_b")
    if plan.num_units and pgt.WindowPlan(offs, W, S).num_units:  # direct kernel implements the same order when gw == 32
        same = all(direct[k].tobytes() == tiled[k].tobytes() for k in tiled)
        assert same or min(W, S) < 256, "tiled and direct level 1 differ"
    for k in tiled:
        assert np.concatenate([p[k] for p in parts]).tobytes() == tiled[k].tobytes(), "shards " + k
        assert host[k].tobytes() == tiled[k].tobytes(), "host " + k


@pytest.mark.parametrize("W,S,unit", [(1000, 1000, 0), (5000, 1000, 512), (100, 10, 0), (300, 300, 32), (100000, 100000, 4096)])
def test_unit_start_table_gives_the_same_bits(pgt, W, S, unit):
    """Plans of more than 32 segments (device mode; pgt_tune unittable = 1 / 2 turns it off / on) let level 1 read unit starts from a table
    written by k_unit_starts and run the INDIRECT kernels (the ones bp mode uses) instead of deriving every unit from
    its segment record: same units, same lanes, so every statistic must come out bit for bit, through the tiled and
    the direct level-1 kernels, unsharded and sharded."""
    rng = np.random.default_rng(W * 7 + S)
    lengths = rng.integers(1, 700, size=1500).tolist() + [1, 2, W, W + S, 3 * W + 1]
    offs = offsets(lengths)
    n = int(offs[-1])
    a, b = pgt.synth_fst(9, 0, n)
    g = pgt.synth_het(9, 0, n)
    f1, f2, n1, n2 = pgt.synth_dxy(9, 0, n)
    pos = pgt.synth_pos(9, 0, n, offs, 1)
    plan = pgt.WindowPlan(offs, W, S, unit_sites=unit)
    args = (pos, a, b, g, f1, f2, n1, n2)
    try:
        for l1 in (0, 1, 2):
            pgt.tune("level1", l1)
            pgt.tune("unittable", 1)
            ref = npy(pgt.fused_window(plan, *args, minind=5))
            pgt.tune("unittable", 2)
            tab = npy(pgt.fused_window(plan, *args, minind=5))
            for k in ref:
                assert tab[k].tobytes() == ref[k].tobytes(), (l1, k)
            parts = []
            for r in range(4):
                wl, wh, sl, sh = plan.shard(r, 4)
                if wh > wl:
                    parts.append(npy(pgt.fused_window(plan, *[x[sl:sh] for x in args], minind=5, window_range=(wl, wh), site_origin=sl)))
            for k in ref:
                if k != "dxy_global":
                    assert np.concatenate([p[k] for p in parts]).tobytes() == ref[k].tobytes(), (l1, "shards", k)
    finally:
        pgt.tune("level1", 0)
        pgt.tune("unittable", 0)
