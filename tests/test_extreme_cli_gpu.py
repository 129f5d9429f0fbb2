"""Drop-in ihsWindow / xpehhWindow end to end on the GPU: byte-identical stdout / stderr / exit code
with the reference transcripts (and the live binaries in oracle/_ref when present).  There is no
floating-point accumulation on this path, so no %g tie allowance is needed."""
import os

import numpy as np
import pytest

import cli_util as U
import oracle_lib as O
import textfmt as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def write_case(c, d):
    fmt = T.ihs_text if c["tool"] == "ihsWindow" else T.xpehh_text
    text = fmt(c["names"], c["lengths"], c["pos"], c["v_micro"])
    if c["trailing_blank"]:
        text += "\n"
    open(os.path.join(d, "in.norm"), "w").write(text)
    if c["chr_len"] is not None:
        keep = [(nm, L) for nm, L in zip(c["names"], c["chr_len"]) if L]
        open(os.path.join(d, "len.txt"), "w").write(T.sizes_text([k[0] for k in keep], [k[1] for k in keep]))


def test_golden_transcripts_through_the_clis(golden_extreme_cases, tmp_path):
    # every 5th transcript through a fresh CLI process (each pays 1-2 s of CUDA start-up); ALL
    # transcripts go through the library in test_extreme_gpu.py
    for i, c in enumerate(golden_extreme_cases):
        if i % 5:
            continue
        d = tmp_path / f"c{i}"
        d.mkdir()
        write_case(c, str(d))
        rc, out, err = U.run(U.ours(c["tool"]), c["argv"], cwd=str(d))
        assert (rc, out, err) == (c["rc"], c["stdout"], c["stderr"]), (i, c["tool"], c["argv"])


def test_stream_parsing_quirks(tmp_path):
    """Blank lines repeat the previous site, short lines keep the previous score, a non-numeric score
    reads as 0, CRLF line ends, a chromosome name without '_' -- transcript of the reference binary."""
    text = ("chr1_5 5 0.1 1 1 0.5 1.5 0\n"
            "\n"
            "chr1_9 9 0.1 1 1\n"
            "chr1_12 12 0.1 1 1 0.5 nan 0\r\n"
            "chr1_14 14 0.1 1 1 0.5 -2.5 1 extra\n"
            "scaffoldX 3 0.1 1 1 0.5 2.25 1\n"
            "scaffoldX 40\n"
            "   \n")
    (tmp_path / "q.norm").write_text(text)
    want = ("chr1\t1\t10\t1.5\t5\t0\t3\n"
            "chr1\t11\t20\t-2.5\t14\t0.5\t2\n"
            "scaffoldX\t1\t10\t2.25\t3\t1\t1\n"
            "scaffoldX\t11\t20\tNA\tNA\tNA\t0\n"
            "scaffoldX\t21\t30\tNA\tNA\tNA\t0\n"
            "scaffoldX\t31\t40\t2.25\t40\t1\t1\n"
            "scaffoldX\t41\t50\t2.25\t40\t1\t1\n")  # the repeated site sits on the window end: it opens the next window
    args = ["q.norm", "-winsize", 10]
    if O.ref_binary("ihsWindow"):
        assert U.run(O.ref_binary("ihsWindow"), args, cwd=str(tmp_path)) == (0, want, "")
    assert U.run(U.ours("ihsWindow"), args, cwd=str(tmp_path)) == (0, want, "")


def test_site_beyond_chrlen_is_an_error(tmp_path):
    (tmp_path / "b.norm").write_text("chr1_5 5 0.1 1 1 0.5 1.5 0\nchr1_50 50 0.1 1 1 0.5 1.5 0\n")
    (tmp_path / "len.txt").write_text("chr1 30\n")
    rc, out, err = U.run(U.ours("ihsWindow"), ["b.norm", "-winsize", 10, "-chrlen", "len.txt"], cwd=str(tmp_path))
    assert rc == 255 and out == "" and "beyond its chromosome length" in err


needs_ref = pytest.mark.skipif(O.ref_binary("ihsWindow") is None, reason="oracle/_ref not built")


@needs_ref
def test_synthetic_clis_match_live_reference(tmp_path):
    """3e5-site synthetic files (multi-threaded parser path), default 100 kb windows and -chrlen padding."""
    lengths = [150000, 100000, 50000]
    names = ["chr1", "chr2", "chrX"]
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
    n = int(offs[-1])
    pos = O.synth_pos(11, offs, 40)
    v = np.round(O.synth_score(11, 0, n) * 1e6).astype(np.int64)
    chrlen = [int(pos[offs[c + 1] - 1]) + 350000 for c in range(3)]
    (tmp_path / "len.txt").write_text(T.sizes_text(names, chrlen))
    (tmp_path / "s.ihs.norm").write_text(T.ihs_text(names, lengths, pos, v))
    (tmp_path / "s.xp.norm").write_text(T.xpehh_text(names, lengths, pos, v))
    for tool, args in (("ihsWindow", ["s.ihs.norm"]), ("ihsWindow", ["s.ihs.norm", "-winsize", 25000, "-cutoff", 1.5, "-chrlen", "len.txt"]),
                       ("xpehhWindow", ["s.xp.norm", 2]), ("xpehhWindow", ["s.xp.norm", -1.5, "-winsize", 1000, "-chrlen", "len.txt"])):
        r = U.run(O.ref_binary(tool), args, cwd=str(tmp_path))
        g = U.run(U.ours(tool), args, cwd=str(tmp_path))
        assert g == r, (tool, args)
        assert len(r[1].splitlines()) > 50
