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
