"""Closed-form enumeration (pgt_plan_*, host-only entry points of the C ABI) vs the oracle's
operational buffer simulation.  No GPU: only plan functions are called."""
import numpy as np
import pytest

import oracle_lib as O
import textfmt as T
from popgenomicstools_b200 import WindowPlan, PgtError


def oracle_windows_sites(lengths, W, S):
    chr_id = T.expand_chr(lengths)
    n = len(chr_id)
    r = O.fst(chr_id, np.arange(n, dtype=np.uint32), np.zeros(n), np.ones(n), W, S)
    return r["first"], r["last"], r["label"], r["n"]


def random_lengths(rng, W, S, ncontig, maxlen):
    out = []
    for _ in range(ncontig):
        m = rng.integers(0, 4)
        if m == 0:
            out.append(int(W + S * rng.integers(0, 4)))
        elif m == 1:
            out.append(int(rng.integers(1, max(2, W - S + 2))))
        else:
            out.append(int(rng.integers(1, maxlen + 1)))
    return out


@pytest.mark.parametrize("seed", range(4))
def test_site_mode_enumeration_matches_oracle(seed):
    rng = np.random.default_rng(seed)
    for _ in range(150):
        W = int(rng.integers(1, 40))
        S = int(rng.integers(1, W + 1))
        lengths = random_lengths(rng, W, S, int(rng.integers(1, 8)), 120)
        offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
        plan = WindowPlan(offs, W, S, unit_sites=32)
        f, l, lab = plan.windows()
        of, ol, olab, on = oracle_windows_sites(lengths, W, S)
        assert plan.num_windows == len(of), (W, S, lengths)
        assert np.array_equal(f, of) and np.array_equal(l, ol) and np.array_equal(lab, olab), (W, S, lengths)
        assert np.array_equal(l - f + 1, on)


@pytest.mark.parametrize("seed", range(3))
def test_bp_mode_enumeration_matches_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    for _ in range(120):
        W = int(rng.integers(1, 30))
        S = int(rng.integers(1, W + 1))
        chr_len = random_lengths(rng, W, S, int(rng.integers(1, 7)), 90)
        # one site per chromosome is enough to make the chromosome "appear in the data"
        pos = np.array([rng.integers(1, L + 1) for L in chr_len], np.uint32)
        chr_id = np.arange(len(chr_len), dtype=np.uint32)
        n = len(pos)
        r = O.dxy(chr_id, pos, np.full(n, .5), np.full(n, .5), np.ones(n), np.ones(n), 1, W, S, 0, 0, chr_len)
        offs = np.concatenate([[0], np.cumsum(chr_len)]).astype(np.uint64)
        plan = WindowPlan(offs, W, S, mode="bp", unit_sites=32)
        f, l, lab = plan.windows()
        assert plan.num_windows == len(r["first"]), (W, S, chr_len)
        assert np.array_equal(f.astype(np.int64), r["first"]) and np.array_equal(l.astype(np.int64), r["last"])
        assert np.array_equal(lab, r["label"])
        # start/end positions derive from entry indices: pos = entry - off[contig] + 1
        cf = np.searchsorted(offs, f, side="right") - 1
        assert np.array_equal((f - offs[cf] + 1).astype(np.int32), r["start"])
        assert np.array_equal((l - offs[lab] + 1).astype(np.int32), r["end"])


def test_units_tile_every_window_exactly():
    rng = np.random.default_rng(5)
    for _ in range(60):
        W = int(rng.integers(1, 300))
        S = int(rng.integers(1, W + 1))
        u = int(rng.choice([32, 64, 256]))
        lengths = random_lengths(rng, W, S, int(rng.integers(1, 5)), 900)
        offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
        plan = WindowPlan(offs, W, S, unit_sites=u)
        units = [plan.unit(j) for j in range(plan.num_units)]
        # units partition the site axis in order, each <= u sites, none empty
        pos = 0
        for st, ln in units:
            assert st == pos and 1 <= ln <= u
            pos += ln
        assert pos == int(offs[-1])
        f, l, _ = plan.windows()
        for w in range(plan.num_windows):
            fu, cnt = plan.window_units(w)
            assert units[fu][0] == f[w]
            assert units[fu + cnt - 1][0] + units[fu + cnt - 1][1] - 1 == l[w]


def test_scale_checks_from_survey():
    # 1e6 sites, 50000/10000 -> 96 windows, last = sites 950000..999999
    plan = WindowPlan([0, 1000000], 50000, 10000)
    f, l, lab = plan.windows()
    assert plan.num_windows == 96 and f[-1] == 950000 and l[-1] == 999999
    # dxy bp: L=1e7, 20000/5000 -> 1997 windows
    plan = WindowPlan([0, 10000000], 20000, 5000, mode="bp")
    f, l, lab = plan.windows()
    assert plan.num_windows == 1997 and f[-1] + 1 == 9980001 and l[-1] + 1 == 10000000


def test_shards_cover_windows_and_cut_on_window_starts():
    lengths = [241170, 230000, 198765, 50000, 1234, 99999]
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    plan = WindowPlan(offs, 5000, 1000)
    f, l, _ = plan.windows()
    for nsh in (1, 2, 3, 4, 8):
        prev = 0
        for r in range(nsh):
            wl, wh, sl, sh = plan.shard(r, nsh)
            assert wl == prev and wh >= wl
            prev = wh
            if wh > wl:
                assert sl == f[wl] and sh == l[wh - 1] + 1
        assert prev == plan.num_windows


def test_argument_errors():
    with pytest.raises(PgtError):
        WindowPlan([0, 10], 0, 1)
    with pytest.raises(PgtError):
        WindowPlan([0, 10], 5, 0)
    with pytest.raises(PgtError):
        WindowPlan([0, 10], 5, 6)   # S > W: the reference segfaults, we refuse
    with pytest.raises(PgtError):
        WindowPlan([0, 10, 5], 2, 1)


def test_many_small_contigs_match_oracle():
    """Draft-assembly shape: tens of thousands of scaffolds, many shorter than the window."""
    rng = np.random.default_rng(11)
    lengths = rng.integers(1, 60, size=20000).tolist()
    W, S = 25, 5
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    plan = WindowPlan(offs, W, S)
    f, l, lab = plan.windows()
    of, ol, olab, on = oracle_windows_sites(lengths, W, S)
    assert np.array_equal(f, of) and np.array_equal(l, ol) and np.array_equal(lab, olab)
    assert plan.num_segments <= len(lengths)


def test_empty_and_degenerate_inputs():
    assert WindowPlan([0], 5, 1).num_windows == 0              # no contigs
    assert WindowPlan([0, 0, 0], 5, 1).num_windows == 0         # only empty contigs
    p = WindowPlan([0, 0, 7, 7, 9], 3, 3)                       # empty contigs between real ones
    f, l, lab = p.windows()
    of, ol, olab, on = oracle_windows_sites([7, 2], 3, 3)
    assert np.array_equal(f, of) and np.array_equal(l, ol)
    assert lab.tolist() == [1, 1, 1, 3]                         # labels index the ORIGINAL contig list
    assert WindowPlan([0, 3], 10, 1).num_windows == 0           # 3 sites <= W-S at EOF: dropped
    assert WindowPlan([0, 3], 10, 8).num_windows == 1           # 3 sites >  W-S at EOF: printed
    big = WindowPlan([0, 2**40], 2**31, 2**20)                  # 64-bit site axis
    assert big.num_windows == (2**40 - 2**31) // 2**20 + 1
