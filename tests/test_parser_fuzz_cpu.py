"""Differential fuzz of the text parsers against the UNMODIFIED reference binaries, on the CPU:

    fuzzed text --(our CLI, PGT_PACK: parser only, no GPU)--> columns --(oracle)--> rows
    fuzzed text --(reference binary)-------------------------------------------> rows

The GPU path is compared with the oracle on columns elsewhere; this closes the loop for the text
layer: number syntax ("+5", ".5", "5.", "1e-3", "1E2"), whitespace mixes, CRLF, blank lines, short
lines, non-numeric tokens, chromosome names with and without '_'."""
import numpy as np
import pytest

import cli_util as U
import oracle_lib as O
import textfmt as T
from popgenomicstools_b200 import colfile

import os

SCALE = int(os.environ.get("PGT_FUZZ_SCALE", "1"))  # more iterations / other seeds for one-off soak runs
SEED = int(os.environ.get("PGT_FUZZ_SEED", "0"))

needs_ref = pytest.mark.skipif(O.ref_binary("ihsWindow") is None, reason="oracle/_ref not built")

NUMS = ["0", "1", "-1", "+2.5", ".5", "-.25", "5.", "1e-3", "-1E2", "2.000000", "0.000001", "-0.0", "3.14159265358979", "123456.789",
        "1e0", "-7.5e-1", "00.5", "4e1"]
BAD = ["nan", "NA", "inf", "-", "abc", "1..2"]


def ids_of(runs):
    """Name id per run: equal names share an id (the oracle compares names through ids)."""
    seen = {}
    return [seen.setdefault(nm, len(seen)) for nm, _ in runs]


def fuzz_norm_text(rng, tool):
    nfields = 6 if tool == "ihsWindow" else 8
    lines, pos = [], 0
    chrs = ["chr1", "chr2", "scaffold", "chr10"]
    ci = 0
    for _ in range(int(rng.integers(5, 120))):
        if rng.random() < 0.08 and ci + 1 < len(chrs):
            ci += 1
            pos = 0
        pos += int(rng.integers(0, 40))
        r = rng.random()
        if r < 0.06 and lines:
            lines.append("" if rng.random() < 0.5 else "  \t ")  # blank line: the previous site again
            continue
        name = chrs[ci] if chrs[ci] == "scaffold" else f"{chrs[ci]}_{pos}"
        sep = rng.choice(["\t", " ", "  ", " \t"])
        k = nfields
        if r < 0.14:
            k = int(rng.integers(0, nfields))  # short line: later fields keep the previous values
        fields = [str(rng.choice(NUMS)) for _ in range(k)]
        if 0.14 <= r < 0.20 and k:
            fields[int(rng.integers(0, k))] = str(rng.choice(BAD))  # extraction stops here
        line = sep.join([name, str(max(pos, 1))] + fields)
        lines.append(line + ("\r" if rng.random() < 0.05 else ""))
    head = "id\tpos\tgpos\tp1\tihh1\tp2\tihh2\txpehh\tnormxpehh\tcrit\n" if tool == "xpehhWindow" else ""
    return head + "\n".join(lines) + ("\n" if rng.random() < 0.8 else "")


# every test runs single-threaded and with three parser threads on tiny chunks (PGT_PARALLEL_MIN_BYTES=1), so
# that chunk boundaries fall everywhere: rows that inherit values across a chunk boundary, runs of equal
# names split between chunks, blank lines at chunk heads
THREADS = [1, 3]


def penv(threads, **kw):
    e = {"PGT_THREADS": str(threads), "PGT_PARALLEL_MIN_BYTES": "1" if threads > 1 else "1048576"}
    e.update({k: str(v) for k, v in kw.items()})
    return e


@needs_ref
@pytest.mark.parametrize("threads", THREADS)
@pytest.mark.parametrize("tool", ["ihsWindow", "xpehhWindow"])
def test_norm_parsers_agree_with_the_reference(tool, threads, tmp_path):
    rng = np.random.default_rng((2026 if tool == "ihsWindow" else 2027) + SEED)
    checked = 0
    for it in range(150 * SCALE):
        text = fuzz_norm_text(rng, tool)
        (tmp_path / "f.norm").write_text(text)
        W = int(rng.choice([1, 7, 25, 100]))
        cutoff = float(rng.choice([2.0, 0.5, 0.0])) if tool == "ihsWindow" else float(rng.choice([1.0, -1.0, -0.5]))
        args = ["f.norm", "-winsize", W, "-cutoff", cutoff] if tool == "ihsWindow" else ["f.norm", cutoff, "-winsize", W]
        rc, out, err = O.run_ref(tool, args, cwd=tmp_path)
        assert rc == 0
        prc, pout, perr = U.run(U.ours(tool), args, cwd=str(tmp_path), env=penv(threads, PGT_PACK=tmp_path / "f.pgtc"))
        assert (prc, pout) == (0, ""), (it, perr)
        c = colfile.read(tmp_path / "f.pgtc", mmap=False)
        if c["nsites"] == 0:
            continue
        ids = ids_of(c["runs"])
        chr_id = np.repeat(np.asarray(ids, np.uint32), [n for _, n in c["runs"]])
        names = {}
        for (nm, _), i in zip(c["runs"], ids):
            names[i] = nm
        r = O.extreme("ihs" if tool == "ihsWindow" else "xpehh", chr_id, c["columns"]["pos"], c["columns"]["score"], W, cutoff)
        rows = O.extreme_rows(r, [names[i] for i in range(len(names))])
        assert rows == out.splitlines(), (it, tool, args, text)
        checked += 1
    assert checked > 100 * SCALE


@needs_ref
@pytest.mark.parametrize("threads", THREADS)
@pytest.mark.parametrize("tool", ["fstWindow", "hetWindow"])
def test_site_parsers_agree_with_the_reference(tool, threads, tmp_path):
    """Well-formed lines in every number syntax the stream extraction accepts; the first empty line
    ends the input (fstWindow.cpp:125)."""
    rng = np.random.default_rng((77 if tool == "fstWindow" else 78) + SEED)
    for it in range(120 * SCALE):
        lines, pos = [], 0
        chrs = ["ctgA", "ctgB", "c_3"]
        ci = 0
        for _ in range(int(rng.integers(1, 80))):
            if rng.random() < 0.1 and ci + 1 < len(chrs):
                ci += 1
                pos = 0
            pos += int(rng.integers(1, 9))
            sep = rng.choice(["\t", " ", "   ", "\t "])
            if tool == "fstWindow":
                vals = [str(rng.choice(NUMS)), str(rng.choice(NUMS))]
            else:
                vals = [str(rng.choice(["0", "1", "2", "-1", "+1", "-9", "3"]))]
            lines.append(sep.join([chrs[ci], str(pos)] + vals) + ("\r" if rng.random() < 0.05 else ""))
        if rng.random() < 0.2 and len(lines) > 3:
            lines.insert(int(rng.integers(1, len(lines))), "")  # everything after the empty line is ignored
        (tmp_path / "f.txt").write_text("\n".join(lines) + "\n")
        W = int(rng.integers(1, 9))
        S = int(rng.integers(1, W + 1))
        rc, out, err = O.run_ref(tool, ["f.txt", W, S], cwd=tmp_path)
        assert rc == 0
        prc, pout, perr = U.run(U.ours(tool), ["f.txt", W, S], cwd=str(tmp_path), env=penv(threads, PGT_PACK=tmp_path / "f.pgtc"))
        assert (prc, pout) == (0, ""), (it, perr, lines)
        c = colfile.read(tmp_path / "f.pgtc", mmap=False)
        if c["nsites"] == 0:
            assert out == ""
            continue
        ids = ids_of(c["runs"])
        chr_id = np.repeat(np.asarray(ids, np.uint32), [n for _, n in c["runs"]])
        names = {i: nm for (nm, _), i in zip(c["runs"], ids)}
        nl = [names[i] for i in range(len(names))]
        if tool == "fstWindow":
            rows = O.fst_rows(O.fst(chr_id, c["columns"]["pos"], c["columns"]["a"], c["columns"]["b"], W, S), nl)
        else:
            rows = O.het_rows(O.het(chr_id, c["columns"]["pos"], c["columns"]["geno"], W, S), nl)
        assert rows == out.splitlines(), (it, tool, W, S, lines)


@needs_ref
@pytest.mark.parametrize("threads", THREADS)
def test_maf_parser_agrees_with_the_reference(threads, tmp_path):
    """dxyWindow: both populations share the site list (the two-file sync is then the identity), fuzzed
    number syntax and white space in the 7-column MAF lines; -fixedsite 1 windows and the global line."""
    rng = np.random.default_rng(4242 + SEED)
    freqs = ["0.000000", "1.000000", "0.5", ".25", "1e-05", "2.5E-1", "+0.125", "0.333333", "0.999999", "1", "0"]
    for it in range(80 * SCALE):
        names = ["chr1", "chr2", "scaf_7"]
        lines1, lines2 = [], []
        ci, pos = 0, 0
        for _ in range(int(rng.integers(2, 60))):
            if rng.random() < 0.1 and ci + 1 < len(names):
                ci += 1
                pos = 0
            pos += int(rng.integers(1, 30))
            for lines in (lines1, lines2):
                sep = rng.choice(["\t", " ", "  ", "\t "])
                al = rng.choice(["A", "C", "G", "T", "N"], size=3)
                lines.append(sep.join([names[ci], str(pos), al[0], al[1], al[2], str(rng.choice(freqs)), str(int(rng.integers(0, 12)))])
                             + ("\r" if rng.random() < 0.03 else ""))
        head = "chromo\tposition\tmajor\tminor\tref\tknownEM\tnInd\n"
        (tmp_path / "p1.mafs").write_text(head + "\n".join(lines1) + "\n")
        (tmp_path / "p2.mafs").write_text(head + "\n".join(lines2) + "\n")
        W = int(rng.integers(1, 7))
        S = int(rng.integers(1, W + 1))
        minind = int(rng.integers(1, 6))
        args = ["-winsize", W, "-stepsize", S, "-minind", minind, "-fixedsite", 1, "p1.mafs", "p2.mafs"]
        rc, out, err = O.run_ref("dxyWindow", args, cwd=tmp_path)
        assert rc == 0
        prc, pout, perr = U.run(U.ours("dxyWindow"), args, cwd=str(tmp_path),
                                env=penv(threads, PGT_PACK=tmp_path / "p1.pgtc", PGT_PACK2=tmp_path / "p2.pgtc"))
        assert (prc, pout) == (0, ""), (it, perr)
        c1, c2 = colfile.read(tmp_path / "p1.pgtc", mmap=False), colfile.read(tmp_path / "p2.pgtc", mmap=False)
        assert c1["runs"] == c2["runs"] and np.array_equal(c1["columns"]["pos"], c2["columns"]["pos"])
        ids = ids_of(c1["runs"])
        chr_id = np.repeat(np.asarray(ids, np.uint32), [n for _, n in c1["runs"]])
        nm = {i: n for (n, _), i in zip(c1["runs"], ids)}
        r = O.dxy(chr_id, c1["columns"]["pos"], c1["columns"]["freq"], c2["columns"]["freq"], c1["columns"]["nind"],
                  c2["columns"]["nind"], minind, W, S, 1)
        assert O.dxy_rows(r, [nm[i] for i in range(len(nm))]) == out.splitlines(), (it, args, lines1[:3])
        assert O.dxy_global_row(r) == err.strip(), (it, err)


@needs_ref
def test_dxy_two_file_sync_agrees_with_the_reference(tmp_path):
    """The two-file sync state machine (dxyWindow.cpp:315-331: position-only catch-up, silent stop on
    interleaved private sites, chromosomes present in one file only) on random pairs of site lists:
    our synced site list (PGT_PACK_SYNCED, no GPU) -> oracle, against the reference's per-site rows."""
    rng = np.random.default_rng(31337 + SEED)
    head = "chromo\tposition\tmajor\tminor\tref\tknownEM\tnInd\n"
    done = 0
    for it in range(250 * SCALE):
        chroms = [f"c{j}" for j in range(int(rng.integers(1, 4)))]
        rows = [[], []]
        for c in chroms:
            base = np.unique(rng.integers(1, 60, size=int(rng.integers(1, 25))))
            kind = rng.integers(0, 6)
            sets = [base, base]
            if kind == 1:    # pop2 lacks some sites
                sets = [base, base[rng.random(len(base)) < 0.7]]
            elif kind == 2:  # pop1 lacks some sites
                sets = [base[rng.random(len(base)) < 0.7], base]
            elif kind == 3:  # private sites on both sides (the reference stops silently there)
                sets = [base[rng.random(len(base)) < 0.8], base[rng.random(len(base)) < 0.8]]
            elif kind == 4 and len(chroms) > 1:  # chromosome missing from one file
                sets = [base, base[:0]] if rng.random() < 0.5 else [base[:0], base]
            for k in (0, 1):
                for p in sets[k]:
                    rows[k].append((c, int(p), int(rng.integers(0, 1000001)), int(rng.integers(0, 12))))
        if not rows[0] or not rows[1]:
            continue
        for k in (0, 1):
            (tmp_path / f"p{k + 1}.mafs").write_text(head + "".join(f"{c}\t{p}\tA\tC\tA\t{T.micro_str(f)}\t{n}\n" for c, p, f, n in rows[k]))
        minind = int(rng.integers(1, 6))
        args = ["-winsize", 1, "-stepsize", 1, "-minind", minind, "-fixedsite", 1, "p1.mafs", "p2.mafs"]
        rc, out, err = O.run_ref("dxyWindow", args, cwd=tmp_path)
        prc, pout, perr = U.run(U.ours("dxyWindow"), args, cwd=str(tmp_path), env={"PGT_PACK_SYNCED": str(tmp_path / "s.pgtc")})
        if rc != 0:  # "Chromosomes in MAF files differ" on the first line: same message, same code
            assert (prc, perr) == (rc, err), (it, rows)
            continue
        assert (prc, pout, perr) == (0, "", ""), (it, perr)
        c = colfile.read(tmp_path / "s.pgtc", mmap=False)
        if c["nsites"] == 0:
            assert out == "", (it, rows)
            continue
        ids = ids_of(c["runs"])
        chr_id = np.repeat(np.asarray(ids, np.uint32), [n for _, n in c["runs"]])
        nm = {i: n for (n, _), i in zip(c["runs"], ids)}
        cols = c["columns"]
        r = O.dxy(chr_id, cols["pos"], cols["f1"], cols["f2"], cols["n1"], cols["n2"], minind, 1, 1, 1)
        assert O.dxy_rows(r, [nm[i] for i in range(len(nm))]) == out.splitlines(), (it, rows)
        assert O.dxy_global_row(r) == err.strip(), (it, err, rows)
        done += 1
    assert done > 150 * SCALE
