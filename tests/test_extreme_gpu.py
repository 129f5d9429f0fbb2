"""CUDA extreme-score window scan (pgt_scan_extreme: ihsWindow / xpehhWindow hot path) against the
oracle and the reference transcripts, through the C ABI.  Everything is bit-exact: the scores are
only compared and counted, never summed (the one division, nbig / nsites, is a single IEEE op)."""
import numpy as np
import pytest

import oracle_lib as O
import textfmt as T
from test_extreme_oracle import case_columns

pytestmark = pytest.mark.gpu

FIELDS = (("ext_value", "ext"), ("ext_pos", "extpos"), ("nbig", "nbig"), ("nsites", "n"), ("prop", "prop"))


def _pgt():
    import popgenomicstools_b200 as pgt
    return pgt


def run_gpu(mode, pos, lengths, val, W, cutoff, chrlen=None, unit_sites=0, host=False, window_range=None):
    import torch
    pgt = _pgt()
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    pos = np.ascontiguousarray(pos, np.uint32)
    val = np.ascontiguousarray(val, np.float64)
    plan = pgt.ExtremePlan(pos, off, W, chrlen, unit_sites)
    fn = pgt.ihs_window if mode == "ihs" else pgt.xpehh_window
    kw = {}
    origin = 0
    if window_range is not None:
        w = plan.windows()
        xoff = np.append(w["first_site"], len(pos))
        origin = int(xoff[window_range[0]])
        kw = dict(window_range=window_range, site_origin=origin)
    if host:
        out = fn(plan, pos[origin:], val[origin:], cutoff, **kw)
    else:
        out = fn(plan, torch.from_numpy(pos[origin:].view(np.int32)).cuda(), torch.from_numpy(val[origin:]).cuda(), cutoff, **kw)
        torch.cuda.synchronize()
        out = {k: v.cpu().numpy() for k, v in out.items()}
    return plan, out


def assert_equal_to_oracle(out, ref, sl=slice(None)):
    for g, r in FIELDS:
        a, b = out[g], ref[r][sl]
        if a.dtype.kind == "f":
            assert np.array_equal(a, b, equal_nan=True), g
        else:
            assert np.array_equal(a, b), g
    empty = ref["n"][sl] == 0
    assert np.all(out["ext_site"][empty] == np.iinfo(np.uint64).max)


def gpu_rows(plan, out, names):
    w = plan.windows()
    res = dict(label=w["label"], start=w["start"], end=w["end"], ext=out["ext_value"], extpos=out["ext_pos"],
               prop=out["prop"], n=out["nsites"])
    return O.extreme_rows(res, names)


@pytest.fixture(params=[1, 2], ids=["two-level", "small-window-kernel"])
def xsmall(request):
    """Both kernels of the extreme scan: the two-level scheme (unit partials + window combine) and k_xsmall (a thread
    per window over a shared-memory tile, chosen for short windows; forced here whenever the longest window fits)."""
    pgt = _pgt()
    pgt.tune("xsmall", request.param)
    yield request.param
    pgt.tune("xsmall", 0)


@pytest.mark.parametrize("host", [False, True])
def test_golden_transcripts(golden_extreme_cases, host, xsmall):
    for i, c in enumerate(golden_extreme_cases):
        chr_id, pos, val = case_columns(c)
        lengths = list(c["lengths"])
        if c["trailing_blank"]:
            lengths[-1] += 1
        mode = "ihs" if c["tool"] == "ihsWindow" else "xpehh"
        plan, out = run_gpu(mode, pos, lengths, val, c["W"], c["cutoff"], c["chr_len"], host=host)
        assert gpu_rows(plan, out, c["names"]) == c["stdout"].splitlines(), (i, c["argv"])


@pytest.mark.parametrize("unit_sites", [0, 1, 3, 64])
def test_random_vs_oracle(unit_sites, xsmall):
    rng = np.random.default_rng(100 + unit_sites)
    for it in range(40):
        mode = "ihs" if it % 2 == 0 else "xpehh"
        W = int(rng.choice([1, 5, 40, 300, 5000, 100000]))
        ncontig = int(rng.integers(1, 5))
        lengths, pos = [], []
        for _ in range(ncontig):
            k = int(rng.integers(1, 3000))
            gap = int(rng.choice([1, 3, 50, 400]))
            p = np.cumsum(rng.integers(0 if it % 7 == 0 else 1, gap + 1, size=k)) + 1
            lengths.append(k)
            pos.extend(p.tolist())
        n = len(pos)
        val = rng.integers(-3, 4, size=n).astype(np.float64) if it % 3 == 0 else rng.normal(size=n)  # ties: first wins
        if it % 5 == 0:
            val[rng.integers(0, n, size=max(1, n // 50))] = np.nan
            val[0] = np.nan  # NaN on the first site of a window stays the extreme (updateMax at nsites == 0)
        if it % 5 == 1:
            val[rng.integers(0, n, size=3)] = np.inf
            val[rng.integers(0, n, size=3)] = -np.inf
        cutoff = float(rng.choice([2.0, 0.0, 1.0])) if mode == "ihs" else float(rng.choice([1.0, -1.0, 0.0, -0.25]))
        ref = O.extreme(mode, T.expand_chr(lengths), pos, val, W, cutoff)
        plan, out = run_gpu(mode, pos, lengths, val, W, cutoff, unit_sites=unit_sites, host=bool(it % 4 == 3))
        assert plan.num_windows == len(ref["n"])
        assert_equal_to_oracle(out, ref)


def test_long_windows_use_partials_and_warp_combine():
    """Windows of 2..32 units (thread combine) and of > 32 units (warp combine), ties everywhere."""
    rng = np.random.default_rng(3)
    lengths = [200000, 70000, 5]
    pos = np.concatenate([np.arange(1, L + 1) for L in lengths])
    val = rng.integers(-4, 5, size=len(pos)).astype(np.float64)
    for W, U in ((1000000, 64), (10000, 2048), (3000, 64), (50000, 0), (200000, 1)):
        for mode, cutoff in (("ihs", 2.0), ("xpehh", -1.5), ("xpehh", 1.5)):
            ref = O.extreme(mode, T.expand_chr(lengths), pos, val, W, cutoff)
            plan, out = run_gpu(mode, pos, lengths, val, W, cutoff, unit_sites=U)
            assert_equal_to_oracle(out, ref)


def test_shards_equal_whole():
    import torch
    pgt = _pgt()
    rng = np.random.default_rng(8)
    lengths = [40000, 25000, 60000]
    pos = np.concatenate([np.cumsum(rng.integers(1, 30, size=L)) for L in lengths])
    val = rng.normal(size=len(pos))
    W = 5000
    ref = O.extreme("ihs", T.expand_chr(lengths), pos, val, W, 2.0)
    plan, whole = run_gpu("ihs", pos, lengths, val, W, 2.0)
    assert_equal_to_oracle(whole, ref)
    for nsh in (2, 3, 8):
        for r in range(nsh):
            w_lo, w_hi, s_lo, s_hi = plan.shard(r, nsh)
            if w_hi == w_lo:
                continue
            for host in (False, True):
                _, part = run_gpu("ihs", pos, lengths, val, W, 2.0, host=host, window_range=(w_lo, w_hi))
                assert_equal_to_oracle(part, ref, slice(w_lo, w_hi))


def test_synthetic_genome_scale_vs_oracle():
    """2e7 synthetic sites over 5 chromosomes, 100 kb windows, -chrlen padding: device generator ==
    CPU twin, every window equal to the oracle."""
    import torch
    pgt = _pgt()
    lengths = [6000000, 5000000, 4000000, 3000000, 2000000]
    off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    n = int(off[-1])
    pos_d = pgt.synth_pos(7, 0, n, off, 5)
    score_d = pgt.synth_score(7, 0, n)
    pos = pos_d.cpu().numpy()
    score = score_d.cpu().numpy()
    assert np.array_equal(score, O.synth_score(7, 0, n))
    chrlen = [int(pos[off[c + 1] - 1]) + 250000 for c in range(5)]
    plan = pgt.ExtremePlan(pos, off, 100000, chrlen)
    out = pgt.ihs_window(plan, pos_d, score_d, 2.0)
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in out.items()}
    ref = O.extreme("ihs", T.expand_chr(lengths), pos, score, 100000, 2.0, chrlen)
    assert plan.num_windows == len(ref["n"]) > 1000
    assert_equal_to_oracle(out, ref)
    out2 = pgt.xpehh_window(plan, pos, score, -2.0)  # host columns
    ref2 = O.extreme("xpehh", T.expand_chr(lengths), pos, score, 100000, -2.0, chrlen)
    assert_equal_to_oracle(out2, ref2)


def test_errors():
    import torch
    pgt = _pgt()
    pos = np.arange(1, 101, dtype=np.uint32)
    plan = pgt.ExtremePlan(pos, [0, 100], 10)
    with pytest.raises(pgt.PgtError):
        pgt.scan_extreme(plan, 7, 2.0, pos, np.zeros(100))
    with pytest.raises(pgt.PgtError):
        pgt.scan_extreme(plan, 0, 2.0, pos, np.zeros(100), window_range=(5, 200))
    with pytest.raises(TypeError):
        pgt.ihs_window(plan, pos, np.zeros(100, np.float32))


def test_short_windows_many_batches():
    """3e6 sites in windows of ~33 sites (1e5 windows, hundreds of shared-memory tiles per scan, empty windows inside and
    at the chromosome ends, ties, NaN on first sites): the small-window kernel is chosen automatically, equals the oracle
    and the two-level scheme bit for bit, whole, sharded and from host memory."""
    import torch
    pgt = _pgt()
    rng = np.random.default_rng(21)
    lengths = [1_700_000, 900_000, 400_003]
    pos = np.concatenate([np.cumsum(rng.integers(1, 6, size=L)) for L in lengths])  # ~3 bp per site
    n = len(pos)
    val = np.round(rng.normal(size=n), 1)  # ties
    val[rng.integers(0, n, size=2000)] = np.nan
    chrlen = [int(pos[sum(lengths[:c + 1]) - 1]) + 1000 for c in range(3)]  # padded: empty windows at the ends
    W = 100
    ref = O.extreme("ihs", T.expand_chr(lengths), pos, val, W, 1.0, chrlen)
    try:
        pgt.tune("xsmall", 0)
        plan, out = run_gpu("ihs", pos, lengths, val, W, 1.0, chrlen)
        assert plan.num_windows == len(ref["n"]) > 90000
        assert_equal_to_oracle(out, ref)
        pgt.tune("xsmall", 1)
        _, two = run_gpu("ihs", pos, lengths, val, W, 1.0, chrlen)
        for k in out:
            assert np.asarray(out[k]).tobytes() == np.asarray(two[k]).tobytes(), k
        pgt.tune("xsmall", 0)
        _, hst = run_gpu("ihs", pos, lengths, val, W, 1.0, chrlen, host=True)
        assert_equal_to_oracle(hst, ref)
        for r in range(3):
            w_lo, w_hi, s_lo, s_hi = plan.shard(r, 3)
            _, part = run_gpu("ihs", pos, lengths, val, W, 1.0, chrlen, window_range=(w_lo, w_hi))
            assert_equal_to_oracle(part, ref, slice(w_lo, w_hi))
        ref2 = O.extreme("xpehh", T.expand_chr(lengths), pos, val, W, -0.5, chrlen)
        _, x2 = run_gpu("xpehh", pos, lengths, val, W, -0.5, chrlen)
        assert_equal_to_oracle(x2, ref2)
    finally:
        pgt.tune("xsmall", 0)
