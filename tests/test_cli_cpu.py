"""CLI contract that needs no GPU: usage text, argument errors, exit codes -- byte-compared with
the reference binaries when oracle/_ref is built, and with the strings of the reference sources
(/root/reference/fstWindow.cpp:23-67, hetWindow.cpp:20-64, dxyWindow.cpp:34-139) otherwise."""
import os

import pytest

import cli_util as U
import oracle_lib as O


@pytest.mark.parametrize("tool", ["fstWindow", "hetWindow", "dxyWindow"])
def test_usage_text_matches_reference(tool):
    rc, out, err = U.run(U.ours(tool), [])
    assert rc == 0 and err == ""
    assert out.startswith("\n") and out.endswith("\n\n")
    if O.ref_binary(tool):
        assert (rc, out, err) == U.run(O.ref_binary(tool), [])
    if tool == "fstWindow":
        assert "fstWindow [ANGSD fst variance component file] [window size (number sites)] [step size (number sites)]\n" in out
        assert "default window size: 1\ndefault step size: 1\n" in out and "(5) Fst\n" in out
    if tool == "dxyWindow":
        assert "-winsize      INT     Window size in base pairs (0 for global calculation) [0]\n" in out
        # one pop file only -> still the help text, rc 0 (dxyWindow.cpp:67-70,537-538)
        assert U.run(U.ours(tool), ["x.maf"])[0:2] == (0, out)


@pytest.mark.parametrize("tool,msg", [("fstWindow", "Unable to open Fst variance components file"),
                                      ("hetWindow", "Unable to open genotypes file")])
def test_site_tool_argument_errors(tool, msg, tmp_path):
    f = tmp_path / "in.txt"
    f.write_text("chr1 1 0.1 0.2\n" if tool == "fstWindow" else "chr1 1 0\n")
    cases = [(["/nonexistent/file"], f"{msg} /nonexistent/file\n"),
             ([f, 0], "Window size must be a positive integer\n"),
             ([f, "abc"], "Window size must be a positive integer\n")]
    for args, want in cases:
        rc, out, err = U.run(U.ours(tool), args)
        assert (rc, out, err) == (255, "", want), args
        if O.ref_binary(tool):
            assert (rc, out, err) == U.run(O.ref_binary(tool), args), args
    # step <= 0: same message as the reference, but we stop (the reference runs into UB)
    rc, out, err = U.run(U.ours(tool), [f, 5, 0])
    assert (rc, out, err) == (255, "", "Step size must be a positive integer\n")
    # step > window: reference segfaults, we refuse
    rc, out, err = U.run(U.ours(tool), [f, 2, 3])
    assert rc == 255 and out == "" and "Step size" in err


def test_dxy_argument_errors(tmp_path):
    head = "chromo\tposition\tmajor\tminor\tref\tknownEM\tnInd\n"
    p1, p2, sz = tmp_path / "p1.maf", tmp_path / "p2.maf", tmp_path / "sizes.txt"
    p1.write_text(head + "A\t1\tA\tC\tA\t0.1\t5\n")
    p2.write_text(head + "B\t1\tA\tC\tA\t0.1\t5\n")
    sz.write_text("A\t10\n")
    tool = U.ours("dxyWindow")
    ref = O.ref_binary("dxyWindow")
    cases = [
        (["/nonexistent/p1", p2], "Unable to open Pop1 MAF file: /nonexistent/p1\n"),
        ([p1, "/nonexistent/p2"], "Unable to open Pop2 MAF file: /nonexistent/p2\n"),
        (["-bogus", 1, p1, p2], "Unknown command: -bogus\n"),
        (["-minind", 0, p1, p2], "-minind must be at least 1\n"),
        (["-winsize", 5, p1, p2], "Must specify a -stepsize > 0 when -winsize is > 0\n"),
        # no -sizefile in bp mode: the reference's check is dead code (see dxywindow_main.cpp), it goes on
        (["-winsize", 5, "-stepsize", 1, p1, p2], "Chromosomes in MAF files differ\n"),
        (["-sizefile", "/nonexistent/sz", p1, p2], "Unable to open sizefile: /nonexistent/sz\n"),
        (["-winsize", 2, "-stepsize", 1, "-fixedsite", 1, p1, p2], "Chromosomes in MAF files differ\n"),
    ]
    for args, want in cases:
        rc, out, err = U.run(tool, args)
        assert (rc, out, err) == (255, "", want), args
        if ref:
            assert (rc, out, err) == U.run(ref, args), args
    bad = tmp_path / "bad_sizes.txt"
    bad.write_text("A\n")
    rc, out, err = U.run(tool, ["-winsize", 2, "-stepsize", 1, "-sizefile", bad, p1, p1])
    assert (rc, out, err) == (255, "", "Unable to correctly parse chromosome size file\n")
    if ref:
        assert (rc, out, err) == U.run(ref, ["-winsize", 2, "-stepsize", 1, "-sizefile", bad, p1, p1])


def test_cli_fails_loudly_without_gpu(tmp_path):
    """No CPU fallback: with valid input but no usable device the tool must refuse, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    f = tmp_path / "in.txt"
    f.write_text("chr1 1 0.1 0.2\nchr1 2 0.1 0.2\n")
    rc, out, err = U.run(U.ours("fstWindow"), [f, 1, 1])
    assert rc == 255 and out == "" and "CUDA" in err
