"""pgt_xplan_* (host window bookkeeping of ihsWindow / xpehhWindow) against the oracle's
operational restatement and the reference transcripts: bit-exact rows, labels, bounds, membership."""
import numpy as np
import pytest

import oracle_lib as O
import textfmt as T
from test_extreme_oracle import case_columns


def plan_for(pos, lengths, W, chrlen=None, unit_sites=0):
    from popgenomicstools_b200 import ExtremePlan
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    return ExtremePlan(np.asarray(pos, np.uint32), off, W, chrlen, unit_sites)


def check_plan_vs_oracle(plan, chr_id, pos, W, chrlen):
    r = O.extreme("ihs", chr_id, pos, np.zeros(len(pos)), W, 2.0, chrlen)
    w = plan.windows()
    assert plan.num_windows == len(r["n"])
    assert np.array_equal(w["label"], r["label"])
    assert np.array_equal(w["start"], r["start"])
    assert np.array_equal(w["end"], r["end"])
    assert np.array_equal(w["nsites"], r["n"])
    ne = r["n"] > 0
    assert np.array_equal(w["first_site"][ne], r["first"][ne])
    # CSR: every site in exactly one window, in file order
    assert int(w["nsites"].sum()) == len(pos)
    assert np.all(np.diff(w["first_site"].astype(np.int64)) >= 0)


def test_plan_matches_golden_transcripts(golden_extreme_cases):
    for i, c in enumerate(golden_extreme_cases):
        chr_id, pos, _ = case_columns(c)
        lengths = list(c["lengths"])
        if c["trailing_blank"]:
            lengths[-1] += 1
        plan = plan_for(pos, lengths, c["W"], c["chr_len"])
        check_plan_vs_oracle(plan, chr_id, pos, c["W"], c["chr_len"])
        w = plan.windows()
        want = [ln.split("\t") for ln in c["stdout"].splitlines()]
        assert len(want) == plan.num_windows, (i, c["argv"])
        for k, row in enumerate(want):
            assert row[0] == c["names"][w["label"][k]] and int(row[1]) == w["start"][k] and int(row[2]) == w["end"][k] \
                and int(row[6]) == w["nsites"][k], (i, k, row)


def test_plan_random_vs_oracle():
    rng = np.random.default_rng(5)
    for it in range(300):
        W = int(rng.choice([1, 2, 3, 7, 10, 64, 1000]))
        ncontig = int(rng.integers(1, 6))
        lengths, pos, chrlen = [], [], []
        for _ in range(ncontig):
            L = int(rng.integers(max(2, W // 3), 30 * W + 40))
            k = int(rng.integers(1, 120))
            p = rng.integers(1, L + 1, size=k)
            if it % 3 == 0:  # window-end hits
                p = np.concatenate([p, W * rng.integers(1, L // W + 2, size=k // 2 + 1)])
            p = np.sort(p[(p >= 1) & (p <= L)])
            if it % 5 == 0:
                p = np.unique(p)
            if len(p) > 1 and p[-1] == L and p[-2] == L:  # two SNPs on the last base: the reference loops
                p = p[:-1]
            if len(p) == 0:
                p = np.array([1])
            lengths.append(len(p))
            pos.extend(p.tolist())
            chrlen.append(L)
        use_len = None if it % 4 == 0 else chrlen
        chr_id = T.expand_chr(lengths)
        try:
            O.extreme("ihs", chr_id, pos, np.zeros(len(pos)), W, 2.0, use_len)
        except ValueError:
            from popgenomicstools_b200 import PgtError
            with pytest.raises(PgtError):
                plan_for(pos, lengths, W, use_len)
            continue
        plan = plan_for(pos, lengths, W, use_len, unit_sites=int(rng.choice([0, 1, 5, 64])))
        check_plan_vs_oracle(plan, chr_id, np.asarray(pos), W, use_len)


def test_plan_unsorted_positions_follow_the_reference():
    """The reference lumps out-of-order sites into the open window; the plan does the same."""
    rng = np.random.default_rng(9)
    for _ in range(50):
        W = int(rng.choice([5, 20]))
        lengths = [int(rng.integers(1, 60)) for _ in range(3)]
        pos = rng.integers(1, 40 * W, size=sum(lengths))
        chr_id = T.expand_chr(lengths)
        check_plan_vs_oracle(plan_for(pos, lengths, W), chr_id, pos, W, None)


def test_plan_errors_and_units():
    from popgenomicstools_b200 import PgtError
    with pytest.raises(PgtError):
        plan_for([1, 2, 3], [3], 0)
    with pytest.raises(PgtError, match="beyond its chromosome length"):
        plan_for([5, 10, 30], [3], 10, [20])  # site past -chrlen: the reference never terminates
    p = plan_for(np.arange(1, 10001), [10000], 1000, unit_sites=64)
    w = p.windows()
    assert p.num_units == int(((w["nsites"] + 63) // 64).sum())
    lo = [p.shard(r, 4) for r in range(4)]
    assert lo[0][0] == 0 and lo[-1][1] == p.num_windows and lo[0][2] == 0 and lo[-1][3] == 10000
    for a, b in zip(lo, lo[1:]):
        assert a[1] == b[0] and a[3] == b[2]
