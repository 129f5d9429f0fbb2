"""Helpers to run the drop-in CLIs (popgenomicstools_b200/bin) next to the reference binaries."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "popgenomicstools_b200", "bin")


def ours(tool):
    alt = os.environ.get("PGT_TEST_BIN")  # e.g. an ASan/UBSan build of the CLIs (make -C csrc asan)
    if alt:
        return os.path.join(alt, tool)
    p = os.path.join(BIN, tool)
    if not os.access(p, os.X_OK):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "popgenomicstools_b200", "csrc")], check=True)
    return p


def run(exe, args, cwd=None, env=None):
    e = dict(os.environ)
    if env:
        e.update(env)
    p = subprocess.run([exe] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, env=e)
    return p.returncode, p.stdout, p.stderr
