"""fstWindow hot path on the GPU (through the C ABI) vs the oracle and the reference transcripts."""
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


def gpu_fst(pgt, lengths, pos, a, b, W, S, unit_sites=0):
    import torch
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    plan = pgt.WindowPlan(offs, W, S, unit_sites=unit_sites)
    dev = torch.device("cuda:0")
    ta = torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(dev)
    tb = torch.from_numpy(np.ascontiguousarray(b, np.float64)).to(dev)
    tp = torch.from_numpy(np.ascontiguousarray(pos, np.uint32).view(np.int32)).to(dev)
    out = pgt.fst_window(plan, tp, ta, tb)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    return plan, res


def check_against_oracle(res, lengths, pos, a, b, W, S):
    chr_id = T.expand_chr(lengths)
    ref = O.fst(chr_id, pos, a, b, W, S)
    absr = O.fst(chr_id, pos, np.abs(a), np.abs(b), W, S)
    P.assert_exact(res["label"], ref["label"], "label")
    P.assert_exact(res["start_pos"], ref["start"], "start")
    P.assert_exact(res["end_pos"], ref["end"], "end")
    P.assert_exact(res["mid_pos"], ref["mid"], "mid")
    P.assert_exact(res["nsites"], ref["n"], "nsites")
    P.assert_sum_close(res["sum_a"], ref["asum"], absr["asum"], "sum_a")
    P.assert_sum_close(res["sum_b"], ref["bsum"], absr["bsum"], "sum_b")
    # the ratio is an IEEE divide of the two sums on both sides
    with np.errstate(divide="ignore", invalid="ignore"):
        want = np.where(res["sum_b"] != 0.0, res["sum_a"] / res["sum_b"], 0.0)
    P.assert_exact(res["fst"], want, "fst = sum_a/sum_b")
    return ref


def test_golden_transcripts(pgt, golden_cases):
    ties = 0
    ncase = 0
    for c in golden_cases:
        if c["tool"] != "fstWindow":
            continue
        a, b = T.micro_to_f64(c["a_micro"]), T.micro_to_f64(c["b_micro"])
        pos = np.asarray(c["pos"], np.uint32)
        plan, res = gpu_fst(pgt, c["lengths"], pos, a, b, c["W"], c["S"], unit_sites=32)
        check_against_oracle(res, c["lengths"], pos, a, b, c["W"], c["S"])
        rows = O.fst_rows(dict(label=res["label"], start=res["start_pos"], end=res["end_pos"], mid=res["mid_pos"],
                               fst=res["fst"], n=res["nsites"]), c["names"])
        ties += P.rows_match_modulo_ties(rows, c["stdout"].splitlines(), float_cols={4})
        ncase += 1
    assert ncase >= 50
    assert ties <= 2, f"{ties} last-digit %g ties against the reference transcripts"


@pytest.mark.parametrize("W,S,unit", [(50000, 10000, 0), (1000, 100, 0), (1000, 1, 0), (1, 1, 0), (777, 13, 32),
                                      (4096, 4096, 64), (300, 299, 0), (100000, 20000, 0)])
def test_synthetic_vs_oracle(pgt, W, S, unit):
    """C1 shape (1e6 sites, 50000/10000 -> 96 windows) and other (W,S) regimes incl. carry contigs."""
    import torch
    n_total = 1_000_000 if W >= 50000 else 200_000
    if S == 1 and W > 1:
        n_total = 60_000
    # contig 0 has (N-W)%S==0 -> cross-contig carry; the last contig is shorter than W-S when possible
    l0 = W + 3 * S if W + 3 * S < n_total // 2 else max(1, n_total // 3)
    rest = n_total - l0
    lengths = [l0, rest * 2 // 3, rest - rest * 2 // 3 - min(7, rest // 10), min(7, rest // 10)]
    lengths = [x for x in lengths if x > 0]
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    n = int(offs[-1])
    a, b = O.synth_fst(1, 0, n)
    pos = O.synth_pos(1, offs, 1)
    ga, gb = pgt.synth_fst(1, 0, n)
    gp = pgt.synth_pos(1, 0, n, offs, 1)
    assert np.array_equal(ga.cpu().numpy(), a) and np.array_equal(gb.cpu().numpy(), b)
    assert np.array_equal(gp.cpu().numpy(), pos)
    plan = pgt.WindowPlan(offs, W, S, unit_sites=unit)
    out = pgt.fst_window(plan, gp, ga, gb)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    check_against_oracle(res, lengths, pos, a, b, W, S)


def test_c1_exact_shape(pgt):
    """BASELINE config 1: one 1 Mb contig, 50 kb windows / 10 kb step -> 96 windows."""
    import torch
    n, W, S = 1_000_000, 50000, 10000
    offs = np.array([0, n], np.uint64)
    a, b = O.synth_fst(1, 0, n)
    pos = O.synth_pos(1, offs, 1)
    plan, res = gpu_fst(pgt, [n], pos, a, b, W, S)
    assert plan.num_windows == 96
    check_against_oracle(res, [n], pos, a, b, W, S)
    assert (res["start_pos"][-1], res["end_pos"][-1], res["mid_pos"][-1], res["nsites"][-1]) == (950001, 1000000, 975000, 50000)


def test_shards_bit_identical(pgt):
    """Deterministic chunking: windows computed per shard (1/2/4/8 shards, halo = W-S sites)
    are bit-identical to the single-shard scan."""
    import torch
    lengths = [241170, 230000, 198765, 50000, 1234, 99999]
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    n, W, S = int(offs[-1]), 5000, 1000
    plan = pgt.WindowPlan(offs, W, S)
    a, b = pgt.synth_fst(4, 0, n)
    pos = pgt.synth_pos(4, 0, n, offs, 1)
    full = pgt.fst_window(plan, pos, a, b)
    torch.cuda.synchronize()
    full = {k: v.cpu().numpy() for k, v in full.items()}
    for nsh in (2, 4, 8):
        parts = []
        for r in range(nsh):
            wl, wh, sl, sh = plan.shard(r, nsh)
            if wh == wl:
                continue
            # each shard generates ONLY its own site range (as a rank would)
            sa, sb = pgt.synth_fst(4, sl, sh - sl)
            sp = pgt.synth_pos(4, sl, sh - sl, offs, 1)
            o = pgt.fst_window(plan, sp, sa, sb, window_range=(wl, wh), site_origin=sl)
            torch.cuda.synchronize()
            parts.append({k: v.cpu().numpy() for k, v in o.items()})
        for k in full:
            cat = np.concatenate([p[k] for p in parts])
            assert cat.tobytes() == full[k].tobytes(), (nsh, k)


@pytest.mark.parametrize("W,S", [(50000, 10000), (4096, 4096), (300, 299), (2560, 256)])
def test_tiled_and_direct_level1_are_bit_identical(pgt, W, S):
    """Both level-1 kernels implement the same summation order (lane-strided partials + butterfly
    per unit): for long units (group width 32) the TMA-staged tiled kernel and the direct
    warp-per-unit kernel must agree bit for bit, at any tile/CTA placement."""
    import torch
    lengths = [W + 3 * S, 123457, 7, 60001]
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    n = int(offs[-1])
    a, b = pgt.synth_fst(8, 0, n)
    pos = pgt.synth_pos(8, 0, n, offs, 1)
    plan = pgt.WindowPlan(offs, W, S)
    res = {}
    try:
        for mode in (1, 2):
            pgt.tune("level1", mode)
            o = pgt.fst_window(plan, pos, a, b)
            torch.cuda.synchronize()
            res[mode] = {k: v.cpu().numpy() for k, v in o.items()}
    finally:
        pgt.tune("level1", 0)
    for k in res[1]:
        assert res[1][k].tobytes() == res[2][k].tobytes(), k
    # unaligned column pointers (views starting at odd elements) exercise the head/tail path
    a2, b2 = torch.empty(n + 3, dtype=torch.float64, device="cuda"), torch.empty(n + 5, dtype=torch.float64, device="cuda")
    a2[3:].copy_(a)
    b2[5:].copy_(b)
    try:
        pgt.tune("level1", 2)
        o = pgt.fst_window(plan, pos, a2[3:], b2[5:])
        torch.cuda.synchronize()
    finally:
        pgt.tune("level1", 0)
    for k in res[1]:
        assert o[k].cpu().numpy().tobytes() == res[1][k].tobytes(), k
