"""Pins oracle/pgt_oracle_extreme.c (ihsWindow / xpehhWindow restatement) to the reference:
committed transcripts of the unmodified binaries (tests/golden/ref_transcripts_extreme.json,
generator make_golden_extreme.py) and, when oracle/_ref/ is built, the live binaries."""
import os

import numpy as np
import pytest

import oracle_lib as O
import textfmt as T


def case_columns(c):
    """Columns as the reference's line loop sees them (a trailing blank line repeats the last site)."""
    chr_id = T.expand_chr(c["lengths"])
    pos = np.asarray(c["pos"], np.uint32)
    val = T.micro_to_f64(c["v_micro"])
    if c["trailing_blank"]:
        chr_id, pos, val = np.append(chr_id, chr_id[-1]), np.append(pos, pos[-1]), np.append(val, val[-1])
    return chr_id, pos, val


def oracle_rows_for_case(c):
    chr_id, pos, val = case_columns(c)
    mode = "ihs" if c["tool"] == "ihsWindow" else "xpehh"
    r = O.extreme(mode, chr_id, pos, val, c["W"], c["cutoff"], c["chr_len"])
    return O.extreme_rows(r, c["names"])


def test_oracle_matches_golden_transcripts(golden_extreme_cases):
    assert len(golden_extreme_cases) >= 150
    nrows = 0
    for i, c in enumerate(golden_extreme_cases):
        assert c["rc"] == 0
        want = c["stdout"].splitlines()
        assert oracle_rows_for_case(c) == want, (i, c["tool"], c["argv"], c["lengths"], c["pos"])
        nrows += len(want)
    assert nrows > 2000


def test_probed_known_answers():
    """Transcript of the compiled reference on a hand-made input (window-end quirk, first site of a
    later chromosome in window 1, trailing blank line, -chrlen padding)."""
    chr_id = [0, 0, 0, 0, 0, 1, 1, 1]
    pos = [5, 10, 20, 20, 47, 35, 36, 36]
    val = [1.5, -2.5, 2.5, 0.5, 0.7, -0.7, 3.0, 3.0]
    r = O.extreme("ihs", chr_id, pos, val, 10, 2.0)
    assert O.extreme_rows(r, ["chr1", "chr2"]) == [
        "chr1\t1\t10\t1.5\t5\t0\t1", "chr1\t11\t20\t-2.5\t10\t1\t1", "chr1\t21\t30\t2.5\t20\t0.5\t2",
        "chr1\t31\t40\tNA\tNA\tNA\t0", "chr1\t41\t50\t0.7\t47\t0\t1", "chr2\t1\t10\t-0.7\t35\t0\t1",
        "chr2\t11\t20\tNA\tNA\tNA\t0", "chr2\t21\t30\tNA\tNA\tNA\t0", "chr2\t31\t40\t3\t36\t1\t2"]
    r = O.extreme("ihs", chr_id, pos, val, 10, 2.0, [50, 60])
    assert O.extreme_rows(r, ["chr1", "chr2"])[-2:] == ["chr2\t41\t50\tNA\tNA\tNA\t0", "chr2\t51\t60\tNA\tNA\tNA\t0"]
    r = O.extreme("ihs", [], [], [], 10, 2.0)
    assert O.extreme_rows(r, []) == ["\t1\t10\tNA\tNA\tNA\t0"]


needs_ref = pytest.mark.skipif(O.ref_binary("ihsWindow") is None, reason="oracle/_ref not built")


@needs_ref
def test_oracle_matches_live_reference_binaries(tmp_path):
    rng = np.random.default_rng(77)
    for it in range(60):
        tool = "ihsWindow" if it % 2 == 0 else "xpehhWindow"
        W = int(rng.choice([1, 4, 7, 25, 1000]))
        ncontig = int(rng.integers(1, 4))
        names = [f"s{j}" for j in range(ncontig)]
        lengths, pos, chr_len = [], [], []
        for _ in range(ncontig):
            L = int(rng.integers(W, 40 * W + 50))
            k = int(rng.integers(1, 200))
            p = np.sort(rng.integers(1, L, size=k))
            if it % 3 == 0:
                p = np.unique(np.concatenate([p, W * np.arange(1, L // W)]))
            p = p[p < L]
            if len(p) == 0:
                p = np.array([1])
            lengths.append(len(p))
            pos.extend(p.tolist())
            chr_len.append(L)
        v = rng.integers(-5000000, 5000001, size=len(pos))
        use_len = it % 4 != 1
        (tmp_path / "len.txt").write_text(T.sizes_text(names, chr_len))
        if tool == "ihsWindow":
            cutoff = 1.75
            (tmp_path / "in.norm").write_text(T.ihs_text(names, lengths, pos, v))
            argv = ["in.norm", "-winsize", W, "-cutoff", cutoff]
        else:
            cutoff = -1.0 if it % 4 == 1 else 1.0
            (tmp_path / "in.norm").write_text(T.xpehh_text(names, lengths, pos, v))
            argv = ["in.norm", cutoff, "-winsize", W]
        if use_len:
            argv += ["-chrlen", "len.txt"]
        rc, out, err = O.run_ref(tool, argv, cwd=tmp_path)
        assert rc == 0
        r = O.extreme("ihs" if tool == "ihsWindow" else "xpehh", T.expand_chr(lengths), pos, T.micro_to_f64(v), W, cutoff,
                      chr_len if use_len else None)
        assert O.extreme_rows(r, names) == out.splitlines(), (it, tool, argv)
