"""Text encodings of the three reference input formats from integer micro-unit columns.

TEST INFRASTRUCTURE.  Values are carried as integers k (= value * 1e6) and printed with
exactly 6 decimals, so the double the reference parses (libstdc++ num_get -> strtod,
/root/reference/fstWindow.cpp:141) equals np.float64(k) / 1e6 bit for bit.
"""
import numpy as np


def micro_str(k):
    k = int(k)
    s = "-" if k < 0 else ""
    k = abs(k)
    return f"{s}{k // 1000000}.{k % 1000000:06d}"


def micro_to_f64(k):
    return np.asarray(k, dtype=np.int64).astype(np.float64) / 1e6


def expand_chr(lengths):
    return np.repeat(np.arange(len(lengths), dtype=np.uint32), lengths)


def fst_text(names, lengths, pos, a_micro, b_micro):
    chr_id = expand_chr(lengths)
    return "".join(f"{names[c]}\t{p}\t{micro_str(a)}\t{micro_str(b)}\n"
                   for c, p, a, b in zip(chr_id, pos, a_micro, b_micro))


def het_text(names, lengths, pos, geno):
    chr_id = expand_chr(lengths)
    return "".join(f"{names[c]} {p} {g}\n" for c, p, g in zip(chr_id, pos, geno))


def maf_text(names, lengths, pos, f_micro, nind):
    chr_id = expand_chr(lengths)
    head = "chromo\tposition\tmajor\tminor\tref\tknownEM\tnInd\n"
    return head + "".join(f"{names[c]}\t{p}\tA\tC\tA\t{micro_str(f)}\t{n}\n"
                          for c, p, f, n in zip(chr_id, pos, f_micro, nind))


def sizes_text(names, chr_len):
    return "".join(f"{nm}\t{ln}\n" for nm, ln in zip(names, chr_len))


def ihs_text(names, lengths, pos, v_micro):
    """selscan `norm --ihs` output: id pos freq ihh1 ihh0 unstd norm crit (no header; ihsWindow.cpp:101,163)."""
    chr_id = expand_chr(lengths)
    return "".join(f"{names[c]}_{p}\t{p}\t0.25\t0.1\t0.2\t{micro_str(-v)}\t{micro_str(v)}\t{int(abs(v) > 2000000)}\n"
                   for c, p, v in zip(chr_id, pos, v_micro))


def xpehh_text(names, lengths, pos, v_micro):
    """selscan `norm --xpehh` output: id pos gpos p1 ihh1 p2 ihh2 xpehh normxpehh crit, one header line
    (xpehhWindow.cpp:94,110-115,165)."""
    chr_id = expand_chr(lengths)
    head = "id\tpos\tgpos\tp1\tihh1\tp2\tihh2\txpehh\tnormxpehh\tcrit\n"
    return head + "".join(f"{names[c]}_{p}\t{p}\t{p / 1e6:.6f}\t0.5\t0.1\t0.25\t0.2\t{micro_str(v // 2)}\t{micro_str(v)}\t0\n"
                          for c, p, v in zip(chr_id, pos, v_micro))


def bgzf_bytes(data, rng=None, level=6):
    """`data` as a BGZF file (bgzip / htslib, SAM spec 4.1): gzip members of at most 64 KB with a 'BC' extra
    subfield holding the member's size, then the empty end-of-file member.  rng: ragged member sizes."""
    import struct
    import zlib

    def member(chunk):
        co = zlib.compressobj(level, zlib.DEFLATED, -15)
        cd = co.compress(chunk) + co.flush()
        head = b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(cd) + 25)
        return head + cd + struct.pack("<II", zlib.crc32(chunk), len(chunk))

    out, o = [], 0
    while o < len(data):
        k = 65280 if rng is None else int(rng.integers(1, 65281))
        out.append(member(data[o:o + k]))
        o += k
    out.append(member(b""))
    return b"".join(out)
