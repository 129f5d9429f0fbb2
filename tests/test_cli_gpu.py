"""Drop-in CLIs end to end on the GPU: same stdout / stderr / exit code as the reference
binaries (committed transcripts, and the live binaries in oracle/_ref when present)."""
import gzip
import os

import numpy as np
import pytest

import cli_util as U
import oracle_lib as O
import parity as P
import textfmt as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def write_case(c, d):
    names, lengths, pos = c["names"], c["lengths"], c["pos"]
    if c["tool"] == "fstWindow":
        open(os.path.join(d, "in.txt"), "w").write(T.fst_text(names, lengths, pos, c["a_micro"], c["b_micro"]))
    elif c["tool"] == "hetWindow":
        open(os.path.join(d, "in.txt"), "w").write(T.het_text(names, lengths, pos, c["geno"]))
    else:
        open(os.path.join(d, "p1.maf"), "w").write(T.maf_text(names, lengths, pos, c["f1_micro"], c["n1"]))
        open(os.path.join(d, "p2.maf"), "w").write(T.maf_text(names, lengths, pos, c["f2_micro"], c["n2"]))
        if c["chr_len"] is not None:
            open(os.path.join(d, "sizes.txt"), "w").write(T.sizes_text(names, c["chr_len"]))


FLOAT_COL = {"fstWindow": {4}, "hetWindow": {4}, "dxyWindow": {3}}


@pytest.mark.parametrize("mode,stride,env", [
    ("default", 4, {}),
    # the streaming-upload path (device-resident columns through the pinned ring, PGT_MEM_DEVICE scan), forced on these tiny inputs
    ("stream", 11, {"PGT_STREAM_MIN_SITES": "1"}),
    # one process, three shards (device 0 three times: CUDA_VISIBLE_DEVICES is set, so the list is taken literally)
    ("devices", 13, {"PGT_DEVICES": "0,0,0", "CUDA_VISIBLE_DEVICES": "0"}),
])
def test_golden_transcripts_through_the_clis(golden_cases, tmp_path, mode, stride, env):
    ties = 0
    # every 4th transcript through a fresh CLI process (each pays 0.3-3 s of CUDA start-up); ALL
    # transcripts go through the library in test_fst_gpu.py / test_stats_gpu.py
    for i, c in enumerate(golden_cases):
        if i % stride:
            continue
        d = tmp_path / f"c{i}"
        d.mkdir()
        write_case(c, str(d))
        rc, out, err = U.run(U.ours(c["tool"]), c["argv"], cwd=str(d), env=env)
        assert rc == c["rc"], (i, c["argv"], err)
        if c["tool"] == "dxyWindow" and c["W"] == 0:
            ties += P.rows_match_modulo_ties(out.splitlines(), c["stdout"].splitlines(), {0})
            assert err == c["stderr"]
            continue
        ties += P.rows_match_modulo_ties(out.splitlines(), c["stdout"].splitlines(), FLOAT_COL[c["tool"]])
        if c["tool"] == "dxyWindow":
            ties += P.rows_match_modulo_ties(err.splitlines(), c["stderr"].splitlines(), {0})
        else:
            assert err == c["stderr"]
    assert ties <= 3, f"{ties} last-digit %g ties"


needs_ref = pytest.mark.skipif(O.ref_binary("fstWindow") is None, reason="oracle/_ref not built")


@needs_ref
def test_synthetic_clis_match_live_reference(tmp_path):
    names = ["chr1", "chr2", "chr3"]
    offs = np.array([0, 250000, 250000 + 90000, 250000 + 90000 + 31000], np.uint64)
    W, S = 50000, 10000
    p = str(tmp_path / "s.fst")
    O.write_text("fst", p, names, offs, seed=1)
    r = U.run(O.ref_binary("fstWindow"), [p, W, S])
    g = U.run(U.ours("fstWindow"), [p, W, S])
    assert g[0] == r[0] and g[2] == r[2]
    assert P.rows_match_modulo_ties(g[1].splitlines(), r[1].splitlines(), {4}) <= 1
    # defaults (W=S=1): one row per site
    r = U.run(O.ref_binary("fstWindow"), [p])
    g = U.run(U.ours("fstWindow"), [p])
    assert len(g[1].splitlines()) == int(offs[-1]) and g[1] == r[1]
    p = str(tmp_path / "s.het")
    O.write_text("het", p, names, offs, seed=3)
    for args in ([p], [p, 100000, 100000], [p, 100000, 20000]):
        assert U.run(U.ours("hetWindow"), args) == U.run(O.ref_binary("hetWindow"), args), args
    p1, p2, sz = str(tmp_path / "p1.maf"), str(tmp_path / "p2.maf"), str(tmp_path / "sizes.txt")
    O.write_text("maf", p1, names, offs, seed=2, density=10, pop=1)
    O.write_text("maf", p2, names, offs, seed=2, density=10, pop=2)
    open(sz, "w").write(T.sizes_text(names, (np.diff(offs) * 10).astype(np.uint32)))
    with open(p1, "rb") as f, gzip.open(p1 + ".gz", "wb") as gz:
        gz.write(f.read())
    for args in (["-winsize", 20000, "-stepsize", 5000, "-minind", 5, "-sizefile", sz, p1, p2],
                 ["-winsize", 20000, "-stepsize", 5000, "-minind", 5, "-sizefile", sz, "-skip_missing", 1, p1 + ".gz", p2],
                 ["-winsize", 500, "-stepsize", 100, "-minind", 5, "-fixedsite", 1, p1, p2],
                 ["-fixedsite", 1, "-minind", 3, p1, p2]):
        r = U.run(O.ref_binary("dxyWindow"), args)
        g = U.run(U.ours("dxyWindow"), args)
        assert g[0] == r[0], args
        assert P.rows_match_modulo_ties(g[1].splitlines(), r[1].splitlines(), {0, 3}) <= 2, args
        assert P.rows_match_modulo_ties(g[2].splitlines(), r[2].splitlines(), {0}) <= 1, args


@needs_ref
def test_dxy_sync_and_missing_size_quirks(tmp_path):
    """position-only catch-up, silent truncation, and the
    'Unable to determine size' failure after the windows flushed so far."""
    head = "chromo\tposition\tmajor\tminor\tref\tknownEM\tnInd\n"

    def maf(path, rows):
        open(path, "w").write(head + "".join(f"{c}\t{p}\tA\tC\tA\t{f:.6f}\t{n}\n" for c, p, f, n in rows))

    cases = []
    # identical lists
    a = [("A", p, 0.1 * ((p % 7) + 1), 10) for p in (2, 3, 5, 8, 9, 12)] + [("B", p, 0.25, 10) for p in (1, 4, 6)]
    b = [("A", p, 0.05 * ((p % 5) + 1), 2 if p == 5 else 10) for p in (2, 3, 5, 8, 9, 12)] + [("B", p, 0.5, 10) for p in (1, 4, 6)]
    cases.append((a, b))
    # pop2 is a superset / subset of pop1
    cases.append((a, sorted(b + [("A", 4, 0.3, 10), ("A", 10, 0.2, 10)], key=lambda r: (r[0], r[1]))))
    cases.append((sorted(a + [("A", 1, 0.3, 10), ("B", 5, 0.2, 10)], key=lambda r: (r[0], r[1])), b))
    # interleaved private sites: silent truncation
    cases.append(([("A", p, 0.2, 10) for p in (1, 7, 9, 11)], [("A", p, 0.6, 10) for p in (1, 3, 9, 11)]))
    cases.append(([("A", p, 0.2, 10) for p in (1, 3, 9, 11)], [("A", p, 0.6, 10) for p in (1, 7, 9, 11)]))
    # a chromosome only in one file
    cases.append((a, [r for r in b if r[0] == "A"]))
    sizes_full = tmp_path / "sizes.txt"
    sizes_full.write_text("A\t14\nB\t7\n")
    sizes_noB = tmp_path / "sizesA.txt"
    sizes_noB.write_text("A\t14\n")
    sizes_noA = tmp_path / "sizesB.txt"
    sizes_noA.write_text("B\t7\n")
    for i, (r1, r2) in enumerate(cases):
        p1, p2 = str(tmp_path / f"p1_{i}.maf"), str(tmp_path / f"p2_{i}.maf")
        maf(p1, r1)
        maf(p2, r2)
        for args in (["-winsize", 4, "-stepsize", 2, "-minind", 5, "-sizefile", sizes_full],
                     ["-winsize", 3, "-stepsize", 1, "-minind", 5, "-fixedsite", 1],
                     ["-winsize", 1, "-stepsize", 1, "-fixedsite", 1],
                     ["-fixedsite", 1],
                     ["-winsize", 4, "-stepsize", 2, "-sizefile", sizes_noB],
                     ["-winsize", 3, "-stepsize", 3, "-sizefile", sizes_noA],
                     ["-winsize", 2, "-stepsize", 1]):
            full = [str(x) for x in args] + [p1, p2]
            r = U.run(O.ref_binary("dxyWindow"), full)
            g = U.run(U.ours("dxyWindow"), full)
            assert g[0] == r[0], (i, full, g, r)
            assert P.rows_match_modulo_ties(g[1].splitlines(), r[1].splitlines(), {0, 3}) <= 1, (i, full, g[1], r[1])
            assert P.rows_match_modulo_ties(g[2].splitlines(), r[2].splitlines(), {0}) <= 1, (i, full, g[2], r[2])


def test_timing_report_is_opt_in(tmp_path):
    p = str(tmp_path / "s.fst")
    O.write_text("fst", p, ["c"], np.array([0, 5000], np.uint64), seed=1)
    rc, out, err = U.run(U.ours("fstWindow"), [p, 100, 50])
    assert rc == 0 and err == ""
    rc, out2, err = U.run(U.ours("fstWindow"), [p, 100, 50], env={"PGT_TIMING": "1"})
    import json
    t = json.loads(err)
    assert out2 == out and t["sites"] == 5000 and t["windows"] == len(out.splitlines())
    assert {"parse_ms", "scan_ms", "format_ms"} <= set(t)


def test_columnar_cache_gives_the_same_rows_as_the_text(tmp_path):
    """PGT_PACK -> .pgtc -> the tool on the cache prints byte for byte what it prints on the text
    (fst, het, dxy in both window modes, ihs, xpehh)."""
    names = ["chr1", "chr2", "chr3"]
    offs = np.array([0, 120000, 170000, 171000], np.uint64)
    d = str(tmp_path)
    for kind, tool, args in (("fst", "fstWindow", [5000, 1000]), ("het", "hetWindow", [4096, 512])):
        O.write_text(kind, os.path.join(d, f"s.{kind}"), names, offs, seed=2, density=3)
        assert U.run(U.ours(tool), [f"s.{kind}"] + args, cwd=d, env={"PGT_PACK": os.path.join(d, f"s.{kind}.pgtc")})[0] == 0
        t = U.run(U.ours(tool), [f"s.{kind}"] + args, cwd=d)
        c = U.run(U.ours(tool), [f"s.{kind}.pgtc"] + args, cwd=d)
        assert t[0] == 0 and len(t[1].splitlines()) > 100 and c == t, tool
    for pop in (1, 2):
        O.write_text("maf", os.path.join(d, f"p{pop}.mafs"), names, offs, seed=2, density=3, pop=pop)
    chr_len = [int(offs[i + 1] - offs[i]) * 3 + 5 for i in range(3)]
    open(os.path.join(d, "sizes.txt"), "w").write(T.sizes_text(names, chr_len))
    assert U.run(U.ours("dxyWindow"), ["-fixedsite", 1, "p1.mafs", "p2.mafs"], cwd=d,
                 env={"PGT_PACK": os.path.join(d, "p1.pgtc"), "PGT_PACK2": os.path.join(d, "p2.pgtc")})[0] == 0
    for opts in (["-winsize", 2000, "-stepsize", 500, "-minind", 5, "-fixedsite", 1],
                 ["-winsize", 20000, "-stepsize", 5000, "-minind", 5, "-sizefile", "sizes.txt"]):
        t = U.run(U.ours("dxyWindow"), opts + ["p1.mafs", "p2.mafs"], cwd=d)
        c = U.run(U.ours("dxyWindow"), opts + ["p1.pgtc", "p2.pgtc"], cwd=d)
        m = U.run(U.ours("dxyWindow"), opts + ["p1.pgtc", "p2.mafs"], cwd=d)  # cache and text can be mixed
        assert t[0] == 0 and len(t[1].splitlines()) > 50 and c == t and m == t, opts
    lengths = np.diff(offs).astype(int).tolist()
    pos = O.synth_pos(2, offs, 40)
    v = np.round(O.synth_score(2, 0, int(offs[-1])) * 1e6).astype(np.int64)
    open(os.path.join(d, "i.norm"), "w").write(T.ihs_text(names, lengths, pos, v))
    open(os.path.join(d, "x.norm"), "w").write(T.xpehh_text(names, lengths, pos, v))
    for tool, f, args in (("ihsWindow", "i", ["-winsize", 50000]), ("xpehhWindow", "x", [-1.5, "-winsize", 50000])):
        assert U.run(U.ours(tool), [f"{f}.norm"] + args, cwd=d, env={"PGT_PACK": os.path.join(d, f"{f}.pgtc")})[0] == 0
        t = U.run(U.ours(tool), [f"{f}.norm"] + args, cwd=d)
        c = U.run(U.ours(tool), [f"{f}.pgtc"] + args, cwd=d)
        assert t[0] == 0 and len(t[1].splitlines()) > 50 and c == t, tool


def test_streaming_upload_gives_the_same_rows(tmp_path):
    """Large single-GPU inputs take the streaming path (pgt_uploader: rows of the finished parse prefix / the cache
    file's column blocks go through a persistent pinned ring into device-resident columns, the scan then runs in
    PGT_MEM_DEVICE).  PGT_STREAM_MIN_SITES=1 forces it on small files: stdout and stderr must be what the
    host-memory path (PGT_STREAM=0) prints, for text and cache inputs, single- and multi-chunk parsing."""
    import json
    names = ["chr1", "chr2", "chr3"]
    offs = np.array([0, 200000, 290000, 291500], np.uint64)
    d = str(tmp_path)
    stream = {"PGT_STREAM_MIN_SITES": "1"}
    for kind, tool, argsets in (("fst", "fstWindow", ([5000, 1000], [1000, 1])), ("het", "hetWindow", ([4096, 512], []))):
        O.write_text(kind, os.path.join(d, f"s.{kind}"), names, offs, seed=4, density=3)
        assert U.run(U.ours(tool), [f"s.{kind}"], cwd=d, env={"PGT_PACK": os.path.join(d, f"s.{kind}.pgtc")})[0] == 0
        for args in argsets:
            base = U.run(U.ours(tool), [f"s.{kind}"] + args, cwd=d, env={"PGT_STREAM": "0"})
            assert base[0] == 0 and base[2] == "" and len(base[1].splitlines()) > 50
            for src, extra in ((f"s.{kind}", {"PGT_PARALLEL_MIN_BYTES": "1", "PGT_THREADS": "5"}), (f"s.{kind}.pgtc", {})):
                got = U.run(U.ours(tool), [src] + args, cwd=d, env=dict(stream, **extra))
                assert got == base, (tool, args, src, extra, got[2][:200])
        rc, out, err = U.run(U.ours(tool), [f"s.{kind}.pgtc"] + list(argsets[0]), cwd=d, env=dict(stream, PGT_TIMING="1"))
        t = json.loads(err.strip().splitlines()[-1])
        assert rc == 0 and t["mode"] == "stream" and t["upload_bytes"] == int(offs[-1]) * (20 if kind == "fst" else 5), t
        rc, out, err = U.run(U.ours(tool), [f"s.{kind}"] + list(argsets[0]), cwd=d, env={"PGT_STREAM": "0", "PGT_TIMING": "1"})
        assert json.loads(err.strip().splitlines()[-1])["mode"] == "host"
    for pop in (1, 2):
        O.write_text("maf", os.path.join(d, f"p{pop}.mafs"), names, offs, seed=4, density=3, pop=pop)
    chr_len = [int(offs[i + 1] - offs[i]) * 3 + 5 for i in range(3)]
    open(os.path.join(d, "sizes.txt"), "w").write(T.sizes_text(names, chr_len))
    for opts in (["-winsize", 2000, "-stepsize", 500, "-minind", 5, "-fixedsite", 1],
                 ["-winsize", 20000, "-stepsize", 5000, "-minind", 5, "-sizefile", "sizes.txt"],
                 ["-winsize", 1000, "-stepsize", 3, "-minind", 5, "-fixedsite", 1]):
        base = U.run(U.ours("dxyWindow"), opts + ["p1.mafs", "p2.mafs"], cwd=d, env={"PGT_STREAM": "0"})
        got = U.run(U.ours("dxyWindow"), opts + ["p1.mafs", "p2.mafs"], cwd=d, env=stream)
        assert base[0] == 0 and len(base[1].splitlines()) > 50
        # rows identical; the global line (stderr) is summed over units vs over sites in the two modes: same counts, sum to 1e-12
        assert got[0] == 0 and got[1] == base[1], opts
        gb, gg = base[2].split(), got[2].split()
        assert gb[1:] == gg[1:] and abs(float(gb[0]) - float(gg[0])) <= 1e-5 * abs(float(gb[0])), (base[2], got[2])
