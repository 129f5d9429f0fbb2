#!/usr/bin/env python3
"""Generate tests/golden/ref_transcripts.json from the UNMODIFIED reference binaries.

Run in the build container (needs oracle/_ref/, i.e. `make -C oracle ref`, which compiles
/root/reference/{fstWindow,hetWindow,dxyWindow}.cpp in place).  The reference ships no tests
or golden vectors, so these transcripts -- inputs as integer micro-unit columns,
outputs as the binaries' exact stdout/stderr/exit code -- are the committed known answers that
pin oracle/pgt_oracle.c and the CUDA path on machines without /root/reference.

    python tests/golden/make_golden.py            # rewrites ref_transcripts.json (seeded)
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
    p = subprocess.run([os.path.join(REF_DIR, tool)] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True)
    return {"rc": p.returncode, "stdout": p.stdout, "stderr": p.stderr}


def contig_lengths(rng, W, S, ncontig, maxlen):
    """Mix of arbitrary lengths and lengths that trigger the carry quirk ((N-W)%S==0)."""
    out = []
    for _ in range(ncontig):
        mode = rng.integers(0, 4)
        if mode == 0:
            out.append(int(W + S * rng.integers(0, 4)))  # exactly-full buffer at the contig change
        elif mode == 1:
            out.append(int(rng.integers(1, max(2, W - S + 2))))  # short contig (<= W-S+1)
        else:
            out.append(int(rng.integers(1, maxlen + 1)))
    return out


def positions(rng, lengths, max_gap=4):
    pos = []
    for L in lengths:
        pos.extend(np.cumsum(rng.integers(1, max_gap + 1, size=L)).tolist())
    return pos


def make_site_cases(rng, tool, ncases):
    cases = []
    for ci in range(ncases):
        W = int(rng.integers(1, 13))
        S = int(rng.integers(1, W + 1))
        ncontig = int(rng.integers(1, 6))
        lengths = contig_lengths(rng, W, S, ncontig, 40)
        names = [f"ctg{j}" for j in range(ncontig)]
        n = sum(lengths)
        pos = positions(rng, lengths)
        case = {"tool": tool, "W": W, "S": S, "names": names, "lengths": lengths, "pos": pos}
        with tempfile.TemporaryDirectory() as d:
            if tool == "fstWindow":
                a = rng.integers(-50000, 50001, size=n)
                b = rng.integers(0, 200001, size=n)
                if ci % 7 == 0:  # a contig of all-zero b: exercises bsum == 0 -> fst 0
                    b[: lengths[0]] = 0
                case["a_micro"], case["b_micro"] = a.tolist(), b.tolist()
                open(os.path.join(d, "in.txt"), "w").write(textfmt.fst_text(names, lengths, pos, a, b))
            else:
                g = rng.choice([-1, 0, 1, 2, -3, 5], size=n, p=[0.1, 0.5, 0.25, 0.1, 0.025, 0.025])
                case["geno"] = g.tolist()
                open(os.path.join(d, "in.txt"), "w").write(textfmt.het_text(names, lengths, pos, g))
            argv = ["in.txt", W, S] if ci % 5 else (["in.txt"] if (W, S) == (1, 1) else ["in.txt", W, S])
            case["argv"] = [str(x) for x in argv]
            case.update(run(tool, argv, d))
        cases.append(case)
    return cases


def make_dxy_cases(rng, ncases):
    cases = []
    for ci in range(ncases):
        fixedsite = int(ci % 2)
        W = int(rng.integers(1, 13))
        S = int(rng.integers(1, W + 1))
        if ci % 11 == 10:
            W, S, fixedsite = 0, 0, 1  # global mode
        ncontig = int(rng.integers(1, 5))
        names = [f"chr{j}" for j in range(ncontig)]
        if fixedsite:
            lengths = contig_lengths(rng, max(W, 1), max(S, 1), ncontig, 30)
            pos = positions(rng, lengths)
            chr_len = None
        else:
            # bp mode: chromosome lengths in bp include carry (L==W+kS), stale-carry (L<=W-S) and arbitrary
            chr_len = contig_lengths(rng, W, S, ncontig, 60)
            lengths, pos = [], []
            for L in chr_len:
                k = int(rng.integers(1, min(L, 12) + 1))
                p = np.sort(rng.choice(np.arange(1, L + 1), size=k, replace=False))
                lengths.append(k)
                pos.extend(p.tolist())
        n = sum(lengths)
        f1 = rng.integers(0, 1000001, size=n)
        f2 = rng.integers(0, 1000001, size=n)
        n1 = rng.integers(0, 11, size=n)
        n2 = rng.integers(0, 11, size=n)
        minind = int(rng.integers(1, 6))
        skip_missing = int(rng.integers(0, 2))
        case = {"tool": "dxyWindow", "W": W, "S": S, "names": names, "lengths": lengths, "pos": pos,
                "f1_micro": f1.tolist(), "f2_micro": f2.tolist(), "n1": n1.tolist(), "n2": n2.tolist(),
                "minind": minind, "fixedsite": fixedsite, "skip_missing": skip_missing, "chr_len": chr_len}
        with tempfile.TemporaryDirectory() as d:
            open(os.path.join(d, "p1.maf"), "w").write(textfmt.maf_text(names, lengths, pos, f1, n1))
            open(os.path.join(d, "p2.maf"), "w").write(textfmt.maf_text(names, lengths, pos, f2, n2))
            argv = ["-winsize", W, "-stepsize", S, "-minind", minind, "-skip_missing", skip_missing]
            if fixedsite:
                argv += ["-fixedsite", 1]
            else:
                open(os.path.join(d, "sizes.txt"), "w").write(textfmt.sizes_text(names, chr_len))
                argv += ["-sizefile", "sizes.txt"]
            argv += ["p1.maf", "p2.maf"]
            case["argv"] = [str(x) for x in argv]
            case.update(run("dxyWindow", argv, d))
        cases.append(case)
    return cases


def main():
    rng = np.random.default_rng(20261018)
    doc = {
        "generator": "tests/golden/make_golden.py",
        "reference": "tplinderoth/PopGenomicsTools fstWindow.cpp / hetWindow.cpp / dxyWindow.cpp, g++ -O3 -Wall, unmodified",
        "cases": make_site_cases(rng, "fstWindow", 60) + make_site_cases(rng, "hetWindow", 40) + make_dxy_cases(rng, 88),
    }
    out = os.path.join(HERE, "ref_transcripts.json")
    with open(out, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print(out, len(doc["cases"]), "cases", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
