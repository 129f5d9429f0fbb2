"""ctypes binding of oracle/libpgtoracle.so + helpers to run the reference binaries.

TEST INFRASTRUCTURE ONLY -- the product package never imports this module.
The oracle restates /root/reference/{fstWindow,hetWindow,dxyWindow}.cpp over
columnar arrays (see oracle/pgt_oracle.c for file:line citations).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
_LIB = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "libpgtoracle.so")
        if not os.path.exists(path):
            build_oracle()
        _LIB = C.CDLL(path)
        _LIB.pgt_oracle_fst.restype = C.c_int64
        _LIB.pgt_oracle_het.restype = C.c_int64
        _LIB.pgt_oracle_dxy.restype = C.c_int64
        _LIB.pgt_oracle_extreme.restype = C.c_int64
    return _LIB


def ref_binary(name):
    """Path of the compiled unmodified reference binary, or None if not built."""
    p = os.path.join(REF_DIR, name)
    return p if os.access(p, os.X_OK) else None


def _p(arr):
    return None if arr is None else arr.ctypes.data_as(C.c_void_p)


def _c(x, dt):
    return np.ascontiguousarray(x, dtype=dt)


def fst(chr_id, pos, a, b, W, S, count_only=False):
    chr_id, pos, a, b = _c(chr_id, np.uint32), _c(pos, np.uint32), _c(a, np.float64), _c(b, np.float64)
    n = len(pos)
    L = lib()
    if count_only:
        return L.pgt_oracle_fst(_p(chr_id), _p(pos), _p(a), _p(b), C.c_uint64(n), C.c_uint32(W), C.c_uint32(S),
                                C.c_uint64(0), *([None] * 10))
    cap = n // S + 2 * (int(chr_id.max()) + 2 if n else 2) + 8 if S > 0 else 8
    o = dict(label=np.zeros(cap, np.uint32), start=np.zeros(cap, np.uint32), end=np.zeros(cap, np.uint32),
             mid=np.zeros(cap, np.uint32), asum=np.zeros(cap), bsum=np.zeros(cap), fst=np.zeros(cap),
             n=np.zeros(cap, np.uint32), first=np.zeros(cap, np.uint64), last=np.zeros(cap, np.uint64))
    r = L.pgt_oracle_fst(_p(chr_id), _p(pos), _p(a), _p(b), C.c_uint64(n), C.c_uint32(W), C.c_uint32(S),
                         C.c_uint64(cap), _p(o["label"]), _p(o["start"]), _p(o["end"]), _p(o["mid"]),
                         _p(o["asum"]), _p(o["bsum"]), _p(o["fst"]), _p(o["n"]), _p(o["first"]), _p(o["last"]))
    if r < 0:
        raise ValueError(f"pgt_oracle_fst error {r}")
    assert r <= cap, (r, cap)
    return {k: v[:r] for k, v in o.items()}


def het(chr_id, pos, geno, W, S):
    chr_id, pos, geno = _c(chr_id, np.uint32), _c(pos, np.uint32), _c(geno, np.int8)
    n = len(pos)
    cap = n // S + 2 * (int(chr_id.max()) + 2 if n else 2) + 8 if S > 0 else 8
    o = dict(label=np.zeros(cap, np.uint32), start=np.zeros(cap, np.uint32), end=np.zeros(cap, np.uint32),
             mid=np.zeros(cap, np.uint32), nhet=np.zeros(cap, np.uint32), nonmissing=np.zeros(cap, np.uint32),
             h=np.zeros(cap), first=np.zeros(cap, np.uint64), last=np.zeros(cap, np.uint64))
    r = lib().pgt_oracle_het(_p(chr_id), _p(pos), _p(geno), C.c_uint64(n), C.c_uint32(W), C.c_uint32(S),
                             C.c_uint64(cap), _p(o["label"]), _p(o["start"]), _p(o["end"]), _p(o["mid"]),
                             _p(o["nhet"]), _p(o["nonmissing"]), _p(o["h"]), _p(o["first"]), _p(o["last"]))
    if r < 0:
        raise ValueError(f"pgt_oracle_het error {r}")
    assert r <= cap
    return {k: v[:r] for k, v in o.items()}


def dxy(chr_id, pos, f1, f2, n1, n2, minind, W, S, fixedsite, skip_missing=0, chr_len=None):
    chr_id, pos = _c(chr_id, np.uint32), _c(pos, np.uint32)
    f1, f2 = _c(f1, np.float64), _c(f2, np.float64)
    n1, n2 = _c(n1, np.int32), _c(n2, np.int32)
    n = len(pos)
    if chr_len is None:
        chr_len = np.zeros(0, np.uint32)
    chr_len = _c(chr_len, np.uint32)
    if W == 0:
        cap = 1
    elif fixedsite:
        cap = n // S + 2 * (int(chr_id.max()) + 2) + 8
    else:
        cap = (int(chr_len.astype(np.uint64).sum()) + n) // S + 2 * len(chr_len) + 8
    o = dict(label=np.zeros(cap, np.uint32), start=np.zeros(cap, np.int32), end=np.zeros(cap, np.int32),
             dxy=np.zeros(cap), neff=np.zeros(cap, np.uint32), nskip=np.zeros(cap, np.uint32),
             first=np.zeros(cap, np.int64), last=np.zeros(cap, np.int64))
    glob = np.zeros(3)
    before = C.c_uint64(0)
    r = lib().pgt_oracle_dxy(_p(chr_id), _p(pos), _p(f1), _p(f2), _p(n1), _p(n2), C.c_uint64(n), C.c_int(minind),
                             C.c_uint32(W), C.c_uint32(S), C.c_int(fixedsite), C.c_int(skip_missing), _p(chr_len),
                             C.c_uint32(len(chr_len)), C.c_uint64(cap), _p(o["label"]), _p(o["start"]),
                             _p(o["end"]), _p(o["dxy"]), _p(o["neff"]), _p(o["nskip"]), _p(o["first"]),
                             _p(o["last"]), _p(glob), C.byref(before))
    nrow = r if r >= 0 else before.value
    assert nrow <= cap
    res = {k: v[:nrow] for k, v in o.items()}
    res["global"] = glob
    res["rc"] = r if r < 0 else 0
    return res


def extreme(mode, chr_id, pos, val, W, cutoff, chrlen=None):
    """mode 'ihs' | 'xpehh' (oracle/pgt_oracle_extreme.c).  chrlen: per name id, 0 = unknown."""
    chr_id, pos, val = _c(chr_id, np.uint32), _c(pos, np.uint32), _c(val, np.float64)
    n = len(pos)
    chrlen = _c(chrlen if chrlen is not None else [], np.uint32)
    m = 0 if mode == "ihs" else 1
    args = [C.c_int(m), _p(chr_id), _p(pos), _p(val), C.c_uint64(n), _p(chrlen), C.c_uint32(len(chrlen)),
            C.c_uint32(W), C.c_double(cutoff)]
    cap = lib().pgt_oracle_extreme(*args, C.c_uint64(0), *([None] * 9))
    if cap < 0:
        raise ValueError(f"pgt_oracle_extreme error {cap}")
    o = dict(label=np.zeros(cap, np.uint32), start=np.zeros(cap, np.uint32), end=np.zeros(cap, np.uint32),
             ext=np.zeros(cap), extpos=np.zeros(cap, np.uint32), nbig=np.zeros(cap, np.uint32), prop=np.zeros(cap),
             n=np.zeros(cap, np.uint32), first=np.zeros(cap, np.uint64))
    r = lib().pgt_oracle_extreme(*args, C.c_uint64(cap), _p(o["label"]), _p(o["start"]), _p(o["end"]), _p(o["ext"]),
                                 _p(o["extpos"]), _p(o["nbig"]), _p(o["prop"]), _p(o["n"]), _p(o["first"]))
    assert r == cap
    return o


def extreme_rows(res, names):
    """printWindow of ihsWindow.cpp:75-84 / xpehhWindow.cpp:65-74."""
    rows = []
    for i in range(len(res["n"])):
        lab = res["label"][i]
        head = [names[lab] if lab != 0xFFFFFFFF else "", str(res["start"][i]), str(res["end"][i])]
        if res["n"][i] > 0:
            rows.append("\t".join(head + [g6(res["ext"][i]), str(res["extpos"][i]), g6(res["prop"][i]), str(res["n"][i])]))
        else:
            rows.append("\t".join(head + ["NA", "NA", "NA", "0"]))
    return rows


# ---- synthetic generator (CPU twin) ----------------------------------------------------

def synth_fst(seed, site0, n):
    a, b = np.empty(n), np.empty(n)
    lib().pgt_oracle_synth_fst(C.c_uint64(seed), C.c_uint64(site0), C.c_uint64(n), _p(a), _p(b))
    return a, b


def synth_het(seed, site0, n):
    g = np.empty(n, np.int8)
    lib().pgt_oracle_synth_het(C.c_uint64(seed), C.c_uint64(site0), C.c_uint64(n), _p(g))
    return g


def synth_dxy(seed, site0, n):
    f1, f2 = np.empty(n), np.empty(n)
    n1, n2 = np.empty(n, np.int32), np.empty(n, np.int32)
    lib().pgt_oracle_synth_dxy(C.c_uint64(seed), C.c_uint64(site0), C.c_uint64(n), _p(f1), _p(f2), _p(n1), _p(n2))
    return f1, f2, n1, n2


def synth_score(seed, site0, n):
    s = np.empty(n)
    lib().pgt_oracle_synth_score(C.c_uint64(seed), C.c_uint64(site0), C.c_uint64(n), _p(s))
    return s


def synth_pos(seed, contig_offsets, density):
    """Positions for all sites of contigs given by offsets[ncontig+1]."""
    n = int(contig_offsets[-1])
    pos = np.empty(n, np.uint32)
    for c in range(len(contig_offsets) - 1):
        lo, hi = int(contig_offsets[c]), int(contig_offsets[c + 1])
        if hi > lo:
            sub = np.empty(hi - lo, np.uint32)
            lib().pgt_oracle_synth_pos(C.c_uint64(seed), C.c_uint64(lo), C.c_uint64(hi - lo), C.c_uint64(lo),
                                       C.c_uint32(density), _p(sub))
            pos[lo:hi] = sub
    return pos


def write_text(kind, path, names, contig_offsets, seed, density=1, pop=1):
    """Write the synthetic genome as the text format `kind` in {'fst','het','maf'}."""
    L = lib()
    for c, name in enumerate(names):
        lo, hi = int(contig_offsets[c]), int(contig_offsets[c + 1])
        args = [path.encode(), C.c_int(1 if c else 0), name.encode(), C.c_uint64(seed), C.c_uint64(lo),
                C.c_uint64(0), C.c_uint64(hi - lo), C.c_uint32(density)]
        if kind == "fst":
            rc = L.pgt_oracle_write_fst_text(*args)
        elif kind == "het":
            rc = L.pgt_oracle_write_het_text(*args)
        else:
            rc = L.pgt_oracle_write_maf_text(*args, C.c_int(pop))
        if rc != 0:
            raise OSError(f"cannot write {path}")


# ---- formatting like the reference's std::cout (default ostream == printf %g) ------------

def g6(x):
    return "%g" % x


def fst_rows(res, names):
    return ["\t".join([names[res["label"][i]], str(res["start"][i]), str(res["end"][i]), str(res["mid"][i]),
                       g6(res["fst"][i]), str(res["n"][i])]) for i in range(len(res["n"]))]


def het_rows(res, names):
    return ["\t".join([names[res["label"][i]], str(res["start"][i]), str(res["end"][i]), str(res["mid"][i]),
                       g6(res["h"][i]), str(res["nonmissing"][i])]) for i in range(len(res["h"]))]


def dxy_rows(res, names):
    return ["\t".join([names[res["label"][i]], str(res["start"][i]), str(res["end"][i]), g6(res["dxy"][i]),
                       str(res["neff"][i]), str(res["nskip"][i])]) for i in range(len(res["dxy"]))]


def dxy_global_row(res):
    g = res["global"]
    return "\t".join([g6(g[0]), str(int(g[1])), str(int(g[2]))])


def run_ref(name, args, cwd=None):
    """Run a reference binary; returns (rc, stdout, stderr)."""
    exe = ref_binary(name)
    if exe is None:
        raise FileNotFoundError(name)
    p = subprocess.run([exe] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True)
    return p.returncode, p.stdout, p.stderr
