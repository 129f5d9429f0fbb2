#!/usr/bin/env python3
"""Generate tests/golden/ref_transcripts_extreme.json from the UNMODIFIED reference binaries
oracle/_ref/{ihsWindow,xpehhWindow} (`make -C oracle ref`).  Same scheme as make_golden.py:
inputs as integer micro-unit columns, outputs as the binaries' exact stdout/stderr/exit code.

    python tests/golden/make_golden_extreme.py        # rewrites ref_transcripts_extreme.json (seeded)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import textfmt  # noqa: E402
from oracle_lib import REF_DIR  # noqa: E402


def run(tool, args, cwd):
    p = subprocess.run([os.path.join(REF_DIR, tool)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True,
                       timeout=20)
    return {"rc": p.returncode, "stdout": p.stdout, "stderr": p.stderr}


def make_positions(rng, W, nsites, L):
    """Sorted positions in [1, L) with many hits on multiples of W (the window-end quirk),
    occasional duplicates, and long gaps (empty windows)."""
    kind = rng.integers(0, 4)
    if kind == 0:  # dense
        cand = rng.integers(1, L, size=nsites)
    elif kind == 1:  # mostly multiples of W
        cand = np.concatenate([W * rng.integers(1, max(2, L // W + 1), size=nsites), rng.integers(1, L, size=nsites // 2 + 1)])
    elif kind == 2:  # clustered with gaps
        base = rng.integers(1, L, size=max(1, nsites // 4))
        cand = np.concatenate([b + rng.integers(0, 2 * W, size=4) for b in base])
    else:  # consecutive multiples chains
        m0 = int(rng.integers(1, 4))
        cand = np.concatenate([W * np.arange(m0, m0 + nsites // 2 + 1), rng.integers(1, L, size=2)])
    cand = cand[(cand >= 1) & (cand < L)]
    if len(cand) == 0:
        cand = np.array([1])
    if rng.integers(0, 3):
        cand = np.unique(cand)
    cand = np.sort(cand)[:nsites]
    return cand.astype(np.int64).tolist()


def make_cases(rng, tool, ncases):
    cases = []
    for ci in range(ncases):
        W = int(rng.choice([1, 2, 3, 5, 10, 16, 50, 100]))
        ncontig = int(rng.integers(1, 5))
        names = [f"chr{j + 1}" for j in range(ncontig)]
        use_len = int(ci % 3 != 0)
        chr_len, lengths, pos = [], [], []
        for _ in range(ncontig):
            L = int(rng.integers(max(2, W // 2), 12 * W + 30))
            p = make_positions(rng, W, int(rng.integers(1, 25)), L)
            if use_len and rng.integers(0, 4) == 0:
                p.append(L)  # a SNP on the last base of the chromosome: opens a degenerate window
            chr_len.append(L)
            lengths.append(len(p))
            pos.extend(p)
        n = len(pos)
        v = rng.integers(-4000000, 4000001, size=n)
        if ci % 5 == 0:  # ties: first extreme must win
            v = (v // 1000000) * 1000000
        if tool == "ihsWindow":
            cutoff = float(rng.choice([2.0, 0.0, 1.5, 3.25]))
            text = textfmt.ihs_text(names, lengths, pos, v)
            argv = ["in.norm", "-winsize", W] + (["-cutoff", cutoff] if ci % 4 else [])
            if ci % 4 == 0:
                cutoff = 2.0
        else:
            cutoff = float(rng.choice([2.0, -2.0, 0.0, -0.5, 1.25]))
            text = textfmt.xpehh_text(names, lengths, pos, v)
            argv = ["in.norm", cutoff, "-winsize", W]
        if ci % 7 == 3:
            text += "\n"  # trailing blank line: the reference re-counts the previous site
        case = {"tool": tool, "W": W, "names": names, "lengths": lengths, "pos": pos, "v_micro": v.tolist(),
                "cutoff": cutoff, "chr_len": chr_len if use_len else None, "trailing_blank": int(ci % 7 == 3)}
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "in.norm"), "w").write(text)
            if use_len:
                # one chromosome missing from the length file now and then (lenmap miss -> chrlen 0)
                drop = int(rng.integers(0, ncontig)) if (ncontig > 1 and ci % 6 == 1) else -1
                if drop >= 0:
                    case["chr_len"] = [0 if j == drop else L for j, L in enumerate(chr_len)]
                open(os.path.join(d, "len.txt"), "w").write(
                    textfmt.sizes_text([nm for j, nm in enumerate(names) if j != drop], [L for j, L in enumerate(chr_len) if j != drop]))
                argv += ["-chrlen", "len.txt"]
            case["argv"] = [str(x) for x in argv]
            case.update(run(tool, argv, d))
        cases.append(case)
    return cases


def main():
    rng = np.random.default_rng(20261019)
    doc = {
        "generator": "tests/golden/make_golden_extreme.py",
        "reference": "tplinderoth/PopGenomicsTools ihsWindow.cpp / xpehhWindow.cpp, g++ -O3 -Wall, unmodified",
        "cases": make_cases(rng, "ihsWindow", 90) + make_cases(rng, "xpehhWindow", 90),
    }
    out = os.path.join(HERE, "ref_transcripts_extreme.json")
    with open(out, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print(out, len(doc["cases"]), "cases", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
