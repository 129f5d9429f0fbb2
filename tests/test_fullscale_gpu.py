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
