"""The CLIs print doubles with std::to_chars(general, 6) (pgt_cli.h put_g); the reference prints them with
`std::cout << double` == printf("%g") (/root/reference/fstWindow.cpp:88, hetWindow.cpp:87, dxyWindow.cpp:190).
tests/integration/format_check.cpp compares the two on ~2e7 values incl. exact rounding ties and both sides of the fast path's hand-over margin in every decade."""
import os
import subprocess

import cli_util as U


def test_put_g_equals_printf_g(tmp_path):
    exe = str(tmp_path / "format_check")
    tools = os.path.join(U.ROOT, "popgenomicstools_b200", "csrc", "tools")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", tools, os.path.join(U.ROOT, "tests", "integration", "format_check.cpp"),
                    "-o", exe, "-lz", "-pthread"], check=True)
    p = subprocess.run([exe, "100000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert " 0 mismatches" in p.stdout
