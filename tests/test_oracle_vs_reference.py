"""Pins oracle/pgt_oracle.c to the reference: committed transcripts of the unmodified
reference binaries (tests/golden/) and, when oracle/_ref/ is built, the live binaries."""
import os

import numpy as np
import pytest

import oracle_lib as O
import textfmt as T


def oracle_rows_for_case(c):
    names, lengths = c["names"], c["lengths"]
    chr_id = T.expand_chr(lengths)
    pos = np.asarray(c["pos"], np.uint32)
    if c["tool"] == "fstWindow":
        r = O.fst(chr_id, pos, T.micro_to_f64(c["a_micro"]), T.micro_to_f64(c["b_micro"]), c["W"], c["S"])
        return O.fst_rows(r, names), None
    if c["tool"] == "hetWindow":
        g = np.clip(np.asarray(c["geno"]), -128, 127)
        r = O.het(chr_id, pos, g, c["W"], c["S"])
        return O.het_rows(r, names), None
    r = O.dxy(chr_id, pos, T.micro_to_f64(c["f1_micro"]), T.micro_to_f64(c["f2_micro"]), c["n1"], c["n2"],
              c["minind"], c["W"], c["S"], c["fixedsite"], c["skip_missing"], c["chr_len"])
    return O.dxy_rows(r, names), O.dxy_global_row(r)


def test_oracle_matches_golden_transcripts(golden_cases):
    assert len(golden_cases) >= 150
    for i, c in enumerate(golden_cases):
        rows, glob = oracle_rows_for_case(c)
        want = c["stdout"].splitlines()
        if c["tool"] == "dxyWindow":
            if c["W"] == 0:
                assert want == [glob], (i, c["argv"])
                continue
            assert c["stderr"].splitlines() == [glob], (i, c["argv"])
        assert rows == want, (i, c["tool"], c["argv"], c["lengths"])


def test_appendix_b_known_answers():
    """SURVEY.md Appendix B.1/B.2 transcripts (produced by the reference binaries)."""
    lengths = [10, 7, 2]
    names = ["A", "B", "C"]
    chr_id = T.expand_chr(lengths)
    i = np.concatenate([np.arange(1, L + 1) for L in lengths])
    r = O.fst(chr_id, 10 * i, 0.5 * i, i + 1.0, 4, 3)
    assert O.fst_rows(r, names) == [
        "A\t10\t40\t25\t0.357143\t4", "A\t40\t70\t55\t0.423077\t4", "A\t70\t100\t85\t0.447368\t4",
        "B\t100\t30\t65\t0.4\t4", "B\t30\t60\t45\t0.409091\t4", "B\t60\t70\t65\t0.433333\t2",
        "C\t10\t20\t15\t0.3\t2"]
    r = O.fst(chr_id, 10 * i, 0.5 * i, i + 1.0, 4, 2)
    assert len(r["n"]) == 8 and O.fst_rows(r, names)[4] == "B\t90\t20\t55\t0.423077\t4"
    r = O.het([0] * 5 + [1] * 2, [1, 2, 3, 4, 5, 1, 2], [0, 1, -1, 2, 1, -1, -1], 3, 2)
    assert O.het_rows(r, ["A", "B"]) == ["A\t1\t3\t2\t0.5\t2", "A\t3\t5\t4\t0.5\t2", "B\t5\t2\t3\t1\t1"]


def test_appendix_b_dxy():
    pA, pB = [2, 3, 5, 8, 9, 12], [1, 4, 6]
    pos = np.array(pA + pB)
    chr_id = np.array([0] * 6 + [1] * 3)
    f1 = np.array([0.1 * ((p % 7) + 1) for p in pA] + [0.25] * 3)
    f2 = np.array([0.05 * ((p % 5) + 1) for p in pA] + [0.5] * 3)
    n1 = np.full(9, 10)
    n2 = np.array([10, 10, 2, 10, 10, 10, 10, 10, 10])
    r = O.dxy(chr_id, pos, f1, f2, n1, n2, 5, 4, 2, 0, 0, [14, 7])
    assert O.dxy_rows(r, ["A", "B"]) == [
        "A\t1\t4\t0.8\t2\t0", "A\t3\t6\t0.44\t1\t1", "A\t5\t8\t0.32\t1\t1", "A\t7\t10\t0.72\t2\t0",
        "A\t9\t12\t0.97\t2\t0", "A\t11\t14\t0.57\t1\t0", "B\t13\t2\t0.5\t1\t0", "B\t1\t4\t1\t2\t0",
        "B\t3\t6\t1\t2\t0", "B\t5\t7\t0.5\t1\t0"]
    assert O.dxy_global_row(r) == "3.59\t8\t1"
    r = O.dxy(chr_id, pos, f1, f2, n1, n2, 5, 3, 1, 1)
    assert O.dxy_rows(r, ["A", "B"]) == [
        "A\t2\t5\t0.8\t2\t1", "A\t3\t8\t0.76\t2\t1", "A\t5\t9\t0.72\t2\t1", "A\t8\t12\t1.29\t3\t0",
        "B\t9\t1\t1.47\t3\t0", "B\t12\t4\t1.57\t3\t0", "B\t1\t6\t1.5\t3\t0"]
    r = O.dxy(chr_id, pos, f1, f2, n1, n2, 1, 0, 0, 1)
    assert O.dxy_global_row(r) == "4.18\t9\t0"


needs_ref = pytest.mark.skipif(O.ref_binary("fstWindow") is None, reason="oracle/_ref not built")


@needs_ref
def test_oracle_matches_live_reference_on_synthetic(tmp_path):
    """Synthetic-generator text through the reference binaries == oracle on the generator's arrays."""
    names = ["chr1", "chr2", "chr3"]
    offs = np.array([0, 25000, 25000 + 9000, 25000 + 9000 + 3100], np.uint64)  # chr1: (N-W)%S==0 -> carry
    chr_id = T.expand_chr(np.diff(offs).astype(int))
    W, S = 5000, 1000
    # fst
    p = str(tmp_path / "s.fst")
    O.write_text("fst", p, names, offs, seed=1, density=1)
    rc, out, err = O.run_ref("fstWindow", [p, W, S])
    a, b = O.synth_fst(1, 0, int(offs[-1]))
    pos = O.synth_pos(1, offs, 1)
    assert rc == 0 and out.splitlines() == O.fst_rows(O.fst(chr_id, pos, a, b, W, S), names)
    # het
    p = str(tmp_path / "s.het")
    O.write_text("het", p, names, offs, seed=3, density=1)
    rc, out, err = O.run_ref("hetWindow", [p, W, S])
    g = O.synth_het(3, 0, int(offs[-1]))
    assert rc == 0 and out.splitlines() == O.het_rows(O.het(chr_id, pos, g, W, S), names)
    # dxy, sparse sites (1 per 10 bp), bp windows and fixed-site windows
    p1, p2, sz = str(tmp_path / "p1.maf"), str(tmp_path / "p2.maf"), str(tmp_path / "sizes.txt")
    O.write_text("maf", p1, names, offs, seed=2, density=10, pop=1)
    O.write_text("maf", p2, names, offs, seed=2, density=10, pop=2)
    chr_len = (np.diff(offs) * 10).astype(np.uint32)
    open(sz, "w").write(T.sizes_text(names, chr_len))
    f1, f2, n1, n2 = O.synth_dxy(2, 0, int(offs[-1]))
    pos10 = O.synth_pos(2, offs, 10)
    rc, out, err = O.run_ref("dxyWindow", ["-winsize", 20000, "-stepsize", 5000, "-minind", 5, "-sizefile", sz, p1, p2])
    r = O.dxy(chr_id, pos10, f1, f2, n1, n2, 5, 20000, 5000, 0, 0, chr_len)
    assert rc == 0 and out.splitlines() == O.dxy_rows(r, names) and err.splitlines() == [O.dxy_global_row(r)]
    rc, out, err = O.run_ref("dxyWindow", ["-winsize", 500, "-stepsize", 100, "-minind", 5, "-fixedsite", 1, p1, p2])
    r = O.dxy(chr_id, pos10, f1, f2, n1, n2, 5, 500, 100, 1)
    assert rc == 0 and out.splitlines() == O.dxy_rows(r, names) and err.splitlines() == [O.dxy_global_row(r)]


@needs_ref
def test_oracle_matches_live_reference_random(tmp_path):
    """Fresh random cases (beyond the committed transcripts) against the live binaries."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as G
    rng = np.random.default_rng(7)
    cases = G.make_site_cases(rng, "fstWindow", 40) + G.make_site_cases(rng, "hetWindow", 40) + G.make_dxy_cases(rng, 80)
    for i, c in enumerate(cases):
        rows, glob = oracle_rows_for_case(c)
        want = c["stdout"].splitlines()
        if c["tool"] == "dxyWindow" and c["W"] == 0:
            assert want == [glob]
        else:
            assert rows == want, (i, c["argv"], c["lengths"])
