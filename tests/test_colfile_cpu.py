"""Binary columnar cache (.pgtc): the CLIs' text parsers -> PGT_PACK -> file -> numpy reader.  Runs
without a GPU (packing never touches the device), so this is also the CPU-side test of the
multi-threaded text parsers: every parsed value is compared with the generator's exact value."""
import gzip
import os

import numpy as np
import pytest

import cli_util as U
import oracle_lib as O
import textfmt as T
from popgenomicstools_b200 import colfile


def pack(tool, args, out, cwd, out2=None, threads=None):
    env = {"PGT_PACK": str(out)}
    if out2:
        env["PGT_PACK2"] = str(out2)
    if threads:
        env["PGT_THREADS"] = str(threads)
    rc, so, se = U.run(U.ours(tool), args, cwd=str(cwd), env=env)
    assert (rc, so, se) == (0, "", ""), (tool, args, se)


def test_roundtrip_python_writer_reader(tmp_path):
    runs = [("chr1", 5), ("scaffold_2", 3), ("chr1", 2)]
    cols = dict(pos=np.arange(10, dtype=np.uint32), a=np.linspace(-1, 1, 10), b=np.arange(10) / 7.0)
    p = tmp_path / "x.pgtc"
    colfile.write(p, "fst", runs, cols)
    r = colfile.read(p)
    assert r["kind"] == "fst" and r["runs"] == runs and r["nsites"] == 10
    assert list(r["offsets"]) == [0, 5, 8, 10]
    for k in cols:
        assert np.array_equal(r["columns"][k], cols[k])
    assert os.path.getsize(p) % 4096 == 0
    with pytest.raises(ValueError):
        (tmp_path / "bad").write_bytes(b"chr1 1 0.1 0.2\n" * 10)
        colfile.read(tmp_path / "bad")


@pytest.mark.parametrize("threads", [1, 5])
def test_fst_and_het_text_parsers(tmp_path, threads):
    names = ["chr1", "chr2", "chrUn_3"]
    offs = np.array([0, 70000, 100000, 100003], np.uint64)
    n = int(offs[-1])
    for kind, tool in (("fst", "fstWindow"), ("het", "hetWindow")):
        txt = tmp_path / f"s.{kind}"
        O.write_text(kind, str(txt), names, offs, seed=3, density=7)
        pack(tool, [txt.name, 100, 50], tmp_path / f"s.{kind}.pgtc", tmp_path, threads=threads)
        r = colfile.read(tmp_path / f"s.{kind}.pgtc")
        assert r["kind"] == kind and r["runs"] == list(zip(names, np.diff(offs).astype(int).tolist()))
        assert np.array_equal(r["columns"]["pos"], O.synth_pos(3, offs, 7))
        if kind == "fst":
            a, b = O.synth_fst(3, 0, n)
            assert np.array_equal(r["columns"]["a"], a) and np.array_equal(r["columns"]["b"], b)
        else:
            assert np.array_equal(r["columns"]["geno"], O.synth_het(3, 0, n))


def test_parser_edge_cases(tmp_path):
    """Whitespace mix, CRLF, signs, exponents, the first empty line ends the input
    (/root/reference/fstWindow.cpp:125), genotypes outside int8 are clamped."""
    (tmp_path / "e.fst").write_text("A 1 0.5 1e-3\nA\t2\t-0.25\t+2\r\nB  3   1.5e2  0.000001\n\nC 9 9 9\n")
    pack("fstWindow", ["e.fst"], tmp_path / "e.pgtc", tmp_path)
    r = colfile.read(tmp_path / "e.pgtc")
    assert r["runs"] == [("A", 2), ("B", 1)]
    assert r["columns"]["pos"].tolist() == [1, 2, 3]
    assert r["columns"]["a"].tolist() == [0.5, -0.25, 150.0] and r["columns"]["b"].tolist() == [1e-3, 2.0, 1e-6]
    (tmp_path / "e.het").write_text("A 1 0\nA 2 -5\nA 3 1\nA 4 300\nA 5 2\n")
    pack("hetWindow", ["e.het"], tmp_path / "eh.pgtc", tmp_path)
    assert colfile.read(tmp_path / "eh.pgtc")["columns"]["geno"].tolist() == [0, -1, 1, 127, 2]
    # malformed line: reported, nothing written
    (tmp_path / "bad.fst").write_text("A 1 0.5 0.1\nA x 0.5 0.1\n")
    rc, so, se = U.run(U.ours("fstWindow"), ["bad.fst"], cwd=str(tmp_path), env={"PGT_PACK": str(tmp_path / "bad.pgtc")})
    assert rc == 255 and "cannot parse line 2" in se and not (tmp_path / "bad.pgtc").exists()


def test_maf_parser_plain_and_gzip(tmp_path):
    names, lengths = ["chr1", "chr2"], [4000, 2500]
    rng = np.random.default_rng(1)
    pos = np.concatenate([np.cumsum(rng.integers(1, 9, size=L)) for L in lengths])
    f1, f2 = rng.integers(0, 1000001, size=6500), rng.integers(0, 1000001, size=6500)
    n1, n2 = rng.integers(0, 30, size=6500), rng.integers(0, 30, size=6500)
    (tmp_path / "p1.mafs").write_text(T.maf_text(names, lengths, pos, f1, n1))
    with gzip.open(tmp_path / "p2.mafs.gz", "wt") as f:
        f.write(T.maf_text(names, lengths, pos, f2, n2))
    pack("dxyWindow", ["-winsize", 10, "-stepsize", 5, "-fixedsite", 1, "p1.mafs", "p2.mafs.gz"], tmp_path / "p1.pgtc", tmp_path,
         out2=tmp_path / "p2.pgtc")
    for path, f, nn in ((tmp_path / "p1.pgtc", f1, n1), (tmp_path / "p2.pgtc", f2, n2)):
        r = colfile.read(path)
        assert r["kind"] == "maf" and r["runs"] == list(zip(names, lengths))
        assert np.array_equal(r["columns"]["pos"], pos)
        assert np.array_equal(r["columns"]["freq"], T.micro_to_f64(f))
        assert np.array_equal(r["columns"]["nind"], nn)


def test_norm_parsers(tmp_path):
    names, lengths = ["chr1", "chr22"], [3000, 1200]
    rng = np.random.default_rng(2)
    pos = np.concatenate([np.cumsum(rng.integers(1, 500, size=L)) for L in lengths])
    v = rng.integers(-4000000, 4000001, size=4200)
    (tmp_path / "i.norm").write_text(T.ihs_text(names, lengths, pos, v) + "\n")  # trailing blank line repeats the last site
    (tmp_path / "x.norm").write_text(T.xpehh_text(names, lengths, pos, v))
    pack("ihsWindow", ["i.norm"], tmp_path / "i.pgtc", tmp_path)
    pack("xpehhWindow", ["x.norm", 2], tmp_path / "x.pgtc", tmp_path)
    ri, rx = colfile.read(tmp_path / "i.pgtc"), colfile.read(tmp_path / "x.pgtc")
    assert ri["kind"] == rx["kind"] == "score"
    assert rx["runs"] == list(zip(names, lengths)) and ri["runs"] == [("chr1", 3000), ("chr22", 1201)]
    assert np.array_equal(rx["columns"]["pos"], pos) and np.array_equal(rx["columns"]["score"], T.micro_to_f64(v))
    assert np.array_equal(ri["columns"]["pos"], np.append(pos, pos[-1]))
    assert np.array_equal(ri["columns"]["score"], np.append(T.micro_to_f64(v), v[-1] / 1e6))


def test_wrong_kind_and_corrupt_files_are_refused(tmp_path):
    colfile.write(tmp_path / "h.pgtc", "het", [("A", 3)], dict(pos=[1, 2, 3], geno=[0, 1, 2]))
    rc, so, se = U.run(U.ours("fstWindow"), ["h.pgtc", 2, 1], cwd=str(tmp_path), env={"PGT_PACK": str(tmp_path / "o.pgtc")})
    assert rc == 255 and "columnar file of another tool" in se
    blob = (tmp_path / "h.pgtc").read_bytes()
    (tmp_path / "t.pgtc").write_bytes(blob[:4096 + 8])  # truncated column block
    rc, so, se = U.run(U.ours("hetWindow"), ["t.pgtc", 2, 1], cwd=str(tmp_path), env={"PGT_PACK": str(tmp_path / "o.pgtc")})
    assert rc == 255 and "truncated" in se
    # re-packing a columnar file reproduces it byte for byte
    rc, so, se = U.run(U.ours("hetWindow"), ["h.pgtc"], cwd=str(tmp_path), env={"PGT_PACK": str(tmp_path / "h2.pgtc")})
    assert rc == 0 and (tmp_path / "h2.pgtc").read_bytes() == blob


def test_truncated_or_corrupt_gzip_is_an_error_not_a_short_result(tmp_path):
    """The reference aborts with boost's gzip_error on a damaged .mafs.gz (dxyWindow.cpp:256-278); a drop-in
    must not print the rows it managed to inflate and exit 0."""
    names, lengths = ["chr1"], [30000]
    rng = np.random.default_rng(4)
    pos = np.cumsum(rng.integers(1, 9, size=30000))
    f, nn = rng.integers(0, 1000001, size=30000), rng.integers(0, 30, size=30000)
    (tmp_path / "p1.mafs").write_text(T.maf_text(names, lengths, pos, f, nn))
    with gzip.open(tmp_path / "ok.mafs.gz", "wt") as g:
        g.write(T.maf_text(names, lengths, pos, f, nn))
    blob = (tmp_path / "ok.mafs.gz").read_bytes()
    (tmp_path / "cut.mafs.gz").write_bytes(blob[:len(blob) // 2])          # truncated member
    (tmp_path / "notrailer.mafs.gz").write_bytes(blob[:-4])                # CRC/ISIZE trailer cut
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 0xff
    (tmp_path / "flip.mafs.gz").write_bytes(bytes(bad))                    # corrupt deflate data / CRC mismatch
    args = ["-winsize", 10, "-stepsize", 5, "-fixedsite", 1]
    for name in ("cut.mafs.gz", "notrailer.mafs.gz", "flip.mafs.gz"):
        out = tmp_path / (name + ".pgtc")
        rc, so, se = U.run(U.ours("dxyWindow"), args + ["p1.mafs", name], cwd=str(tmp_path),
                           env={"PGT_PACK": str(tmp_path / "a.pgtc"), "PGT_PACK2": str(out)})
        assert rc == 255 and so == "" and "gzip" in se and not out.exists(), (name, rc, se)
    # two concatenated members (bgzip style) are still one valid input
    (tmp_path / "two.mafs.gz").write_bytes(gzip.compress(b"chromo\tposition\tmajor\tminor\tref\tknownEM\tnInd\n") +
                                           gzip.compress(b"chr1\t5\tA\tC\tA\t0.25\t7\n"))
    rc, so, se = U.run(U.ours("dxyWindow"), args + ["two.mafs.gz", "two.mafs.gz"], cwd=str(tmp_path),
                       env={"PGT_PACK": str(tmp_path / "t1.pgtc"), "PGT_PACK2": str(tmp_path / "t2.pgtc")})
    assert (rc, se) == (0, "") and colfile.read(tmp_path / "t2.pgtc")["columns"]["pos"].tolist() == [5]


def test_bgzf_members_inflate_in_parallel_to_the_same_columns(tmp_path):
    """ANGSD writes .mafs.gz through htslib's BGZF: independent <= 64 KB gzip members that carry their own sizes,
    inflated here by several threads straight to their final offsets.  Same columns as the plain file for one
    and many threads and for the serial zlib stream; a damaged member is an error (CRC / size checked per member),
    a file that only looks like BGZF falls back to the serial stream."""
    import gzip as gz
    import struct
    names, lengths = ["chr1", "chr2", "scaf_3"], [150000, 90000, 7]
    n = sum(lengths)
    rng = np.random.default_rng(11)
    pos = np.concatenate([np.cumsum(rng.integers(1, 9, size=L)) for L in lengths])
    f, nn = rng.integers(0, 1000001, size=n), rng.integers(0, 30, size=n)
    text = T.maf_text(names, lengths, pos, f, nn).encode()
    (tmp_path / "p.mafs").write_bytes(text)
    blob = T.bgzf_bytes(text, rng)
    assert gz.decompress(blob) == text and blob.count(b"BC\x02\x00") > 50
    (tmp_path / "p.mafs.gz").write_bytes(blob)
    (tmp_path / "full.mafs.gz").write_bytes(T.bgzf_bytes(text))  # full 65280-byte members, as bgzip writes them
    args = ["-winsize", 10, "-stepsize", 5, "-fixedsite", 1]
    want = None
    for name, env in (("p.mafs", {}), ("p.mafs.gz", {"PGT_THREADS": "1"}), ("p.mafs.gz", {"PGT_THREADS": "7"}),
                      ("full.mafs.gz", {"PGT_THREADS": "3"}), ("p.mafs.gz", {"PGT_BGZF_PARALLEL": "0"})):
        out = tmp_path / "o2.pgtc"
        rc, so, se = U.run(U.ours("dxyWindow"), args + ["p.mafs", name], cwd=str(tmp_path),
                           env=dict(env, PGT_PACK=str(tmp_path / "o1.pgtc"), PGT_PACK2=str(out)))
        assert (rc, se) == (0, ""), (name, env, rc, se)
        got = out.read_bytes()
        want = want or got
        assert got == want, (name, env)
        out.unlink()
    r = colfile.read(tmp_path / "o1.pgtc")
    assert np.array_equal(r["columns"]["pos"], pos) and np.array_equal(r["columns"]["nind"], nn)
    # damage: a flipped byte inside a member's deflate data, a wrong ISIZE, a file cut inside a member
    members = [m.start() for m in __import__("re").finditer(b"\x1f\x8b\x08\x04\0\0\0\0\x00\xff\x06\x00BC\x02\x00", blob)]
    mid = members[len(members) // 2]
    bad = bytearray(blob)
    bad[mid + 18 + 5] ^= 0x55
    (tmp_path / "flip.mafs.gz").write_bytes(bytes(bad))
    bad = bytearray(blob)
    nxt = members[len(members) // 2 + 1]
    bad[nxt - 4:nxt] = struct.pack("<I", struct.unpack("<I", blob[nxt - 4:nxt])[0] + 1)
    (tmp_path / "isize.mafs.gz").write_bytes(bytes(bad))
    (tmp_path / "cut.mafs.gz").write_bytes(blob[:mid + 40])
    for name in ("flip.mafs.gz", "isize.mafs.gz", "cut.mafs.gz"):
        rc, so, se = U.run(U.ours("dxyWindow"), args + ["p.mafs", name], cwd=str(tmp_path),
                           env={"PGT_THREADS": "4", "PGT_PACK": str(tmp_path / "a.pgtc"), "PGT_PACK2": str(tmp_path / "bad.pgtc")})
        assert rc == 255 and so == "" and "gzip" in se and not (tmp_path / "bad.pgtc").exists(), (name, rc, se)
    # a 'BC' subfield that lies about the member size: not BGZF after all, the serial stream still reads it
    lie = bytearray(blob)
    lie[members[3] + 16:members[3] + 18] = struct.pack("<H", 17)
    (tmp_path / "lie.mafs.gz").write_bytes(bytes(lie))
    rc, so, se = U.run(U.ours("dxyWindow"), args + ["p.mafs", "lie.mafs.gz"], cwd=str(tmp_path),
                       env={"PGT_PACK": str(tmp_path / "a.pgtc"), "PGT_PACK2": str(tmp_path / "l2.pgtc")})
    assert (rc, se) == (0, "") and (tmp_path / "l2.pgtc").read_bytes() == want


def test_hostile_pgtc_headers_are_refused(tmp_path):
    """64-bit header fields that wrap when multiplied (nsites ~ 2^61, huge run counts / name blocks) must be
    rejected by the C++ view and by the Python reader, not turned into out-of-bounds column pointers."""
    import struct
    colfile.write(tmp_path / "ok.pgtc", "fst", [("A", 4)], dict(pos=[1, 2, 3, 4], a=[.1, .2, .3, .4], b=[1, 1, 1, 1]))
    blob = bytearray((tmp_path / "ok.pgtc").read_bytes())
    hdr = struct.Struct("<8sIIQIIQQQQ")
    magic, ver, kid, n, nruns, ncols, nbytes, data_off, r0, r1 = hdr.unpack(bytes(blob[:64]))

    def variant(name, **kw):
        f = dict(n=n, nruns=nruns, nbytes=nbytes, data_off=data_off)
        f.update(kw)
        b = bytearray(blob)
        b[:64] = hdr.pack(magic, ver, kid, f["n"], f["nruns"], ncols, f["nbytes"], f["data_off"], r0, r1)
        if "count0" in kw:
            b[64:72] = struct.pack("<Q", kw["count0"])
        (tmp_path / name).write_bytes(bytes(b))
        return name

    cases = [variant("wrap.pgtc", n=(1 << 61) + 4, count0=(1 << 61) + 4),   # nsites * 8 wraps to 32
             variant("wrap2.pgtc", n=1 << 63, count0=1 << 63),
             variant("names.pgtc", nbytes=(1 << 64) - 64 - 8 + 2),          # meta_end wraps
             variant("runs.pgtc", nruns=(1 << 32) - 1),
             variant("off.pgtc", data_off=1 << 40),
             variant("count.pgtc", count0=(1 << 64) - 1)]                    # run counts wrap around to nsites
    for name in cases:
        rc, so, se = U.run(U.ours("fstWindow"), [name, 2, 1], cwd=str(tmp_path), env={"PGT_PACK": str(tmp_path / "o.pgtc")})
        assert rc == 255 and so == "" and ("corrupt" in se or "truncated" in se), (name, rc, se)
        assert not (tmp_path / "o.pgtc").exists()
        with pytest.raises(ValueError):
            colfile.read(tmp_path / name)
