"""CopyPool (csrc/pgt_hostcopy.h): the threaded memcpy behind the experimental pinned staging ring of host mode
(pgt_tune "hoststage").  Exercised on the CPU, plain and under ThreadSanitizer."""
import os
import subprocess

import pytest

import cli_util as U


@pytest.mark.parametrize("san", [[], ["-fsanitize=thread"]])
def test_copypool(san, tmp_path):
    exe = str(tmp_path / "copypool_check")
    csrc = os.path.join(U.ROOT, "popgenomicstools_b200", "csrc")
    subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-pthread", *san, "-I", csrc,
                    os.path.join(U.ROOT, "tests", "integration", "copypool_check.cpp"), "-o", exe], check=True)
    p = subprocess.run([exe], capture_output=True, text=True, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1"))
    assert p.returncode == 0 and "copies ok" in p.stdout, p.stdout + p.stderr
