"""ihsWindow / xpehhWindow CLI contract that needs no GPU: usage text, argument errors, exit codes and
the no-site outputs -- byte-compared with the reference binaries when oracle/_ref is built, and with
the strings of /root/reference/ihsWindow.cpp:16-67, xpehhWindow.cpp:16-73 otherwise."""
import pytest

import cli_util as U
import oracle_lib as O


def both(tool, args, cwd=None):
    got = U.run(U.ours(tool), args, cwd=cwd)
    if O.ref_binary(tool):
        assert got == U.run(O.ref_binary(tool), args, cwd=cwd), (tool, args)
    return got


def test_usage_text():
    rc, out, err = both("ihsWindow", [])
    assert rc == 1 and err == "Must supply iHS input file\n"
    assert "ihsWindow [selscan normalized iHS *.norm file] [options]\n" in out
    assert "-winsize INT Window size (bp) [100000]\n" in out and "|iHS| > cutoff [2]\n" in out
    rc, out, err = both("xpehhWindow", [])
    assert rc == 1 and err == "Must supply XPEHH file and cutoff value\n"
    assert "xpehhWindow <selscan normalized XPEHH *.norm file> <cutoff> [options]\n" in out
    assert both("xpehhWindow", ["only_file.norm"])[0] == 1


def test_argument_errors(tmp_path):
    f = tmp_path / "in.norm"
    f.write_text("chr1_5\t5\t0.25\t0.1\t0.2\t-1.5\t1.5\t0\n")
    assert both("ihsWindow", ["/nonexistent/x"]) == (255, "", "Unable to open iHS file /nonexistent/x\n")
    assert both("xpehhWindow", ["/nonexistent/x", 2]) == (255, "", "Unable to open XPEHH inpt file /nonexistent/x\n")
    assert both("ihsWindow", [f, "-winsize", 0]) == (255, "", "Window size must be a positive integer\n")
    assert both("ihsWindow", [f, "-cutoff", -1]) == (255, "", "|iHS| cutoff must be >= zero\n")
    assert both("ihsWindow", [f, "-chrlen", "/nonexistent/len"]) == (255, "", "Unable to open chromosome length file /nonexistent/len\n")
    assert both("ihsWindow", [f, "-bogus", 1]) == (255, "", "Unknown argument -bogus\n")
    assert both("xpehhWindow", [f, 2, "-cutoff", 1]) == (255, "", "Unknown argument -cutoff\n")
    assert both("xpehhWindow", [f, 2, "-winsize", 0]) == (255, "", "Window size must be a positive integer\n")
    # a flag without its value: the reference reads argv[argc]; we refuse
    rc, out, err = U.run(U.ours("ihsWindow"), [f, "-winsize"])
    assert rc == 255 and out == "" and "-winsize" in err


def test_inputs_without_sites(tmp_path):
    """No site at all: ihsWindow still prints its open window with an empty name (ihsWindow.cpp:180);
    xpehhWindow reports a header-less file (xpehhWindow.cpp:111-115).  No GPU work is involved."""
    e = tmp_path / "empty.norm"
    e.write_text("")
    assert both("ihsWindow", [e, "-winsize", 50]) == (0, "\t1\t50\tNA\tNA\tNA\t0\n", "")
    assert both("xpehhWindow", [e, 2]) == (0, "", "Input XPEHH file had zero sites\n")
    h = tmp_path / "header.norm"
    h.write_text("id\tpos\tgpos\tp1\tihh1\tp2\tihh2\txpehh\tnormxpehh\tcrit\n")
    assert both("xpehhWindow", [h, 2]) == (0, "\t1\t100000\tNA\tNA\tNA\t0\n", "")
    rc, out, err = both("xpehhWindow", [h, 0])
    assert rc == 0 and err.startswith("WARNING: cutoff value of zero")
