"""Reader / writer of the `.pgtc` binary columnar cache (popgenomicstools_b200/csrc/tools/pgt_colfile.h).

A `.pgtc` file holds what the drop-in CLIs' text parsers produce -- the columns libpgtscan reads and
the runs of equal chromosome name -- for one input file of fstWindow (/root/reference/fstWindow.cpp:17-21),
hetWindow (hetWindow.cpp:18), one population MAF of dxyWindow (dxyWindow.cpp:24-32) or a selscan
.norm file of ihsWindow / xpehhWindow (ihsWindow.cpp:119,160).  The CLIs write it with
PGT_PACK=<out.pgtc> and accept it wherever the text file went.  numpy only; no GPU involved.
"""
import struct

import os

import numpy as np

MAGIC = b"PGTCOL\x01\n"
ALIGN = 4096
KINDS = {
    "fst": (1, (("pos", np.uint32), ("a", np.float64), ("b", np.float64))),
    "het": (2, (("pos", np.uint32), ("geno", np.int8))),
    "maf": (3, (("pos", np.uint32), ("freq", np.float64), ("nind", np.int32))),
    "score": (4, (("pos", np.uint32), ("score", np.float64))),
    "dxy": (5, (("pos", np.uint32), ("f1", np.float64), ("f2", np.float64), ("n1", np.int32), ("n2", np.int32))),
}
_BY_ID = {v[0]: (k, v[1]) for k, v in KINDS.items()}
_HEADER = struct.Struct("<8sIIQIIQQQQ")


def _pad(x):
    return (x + ALIGN - 1) // ALIGN * ALIGN


def write(path, kind, runs, columns):
    """runs: [(chromosome name, site count), ...] in file order; columns: dict name -> array."""
    kid, spec = KINDS[kind]
    n = sum(int(c) for _, c in runs)
    names = b"".join(nm.encode() + b"\0" for nm, _ in runs)
    data_off = _pad(_HEADER.size + 8 * len(runs) + len(names))
    with open(path, "wb") as f:
        f.write(_HEADER.pack(MAGIC, 1, kid, n, len(runs), len(spec), len(names), data_off, 0, 0))
        f.write(np.asarray([c for _, c in runs], dtype="<u8").tobytes())
        f.write(names)
        f.write(b"\0" * (data_off - f.tell()))
        for name, dt in spec:
            col = np.ascontiguousarray(columns[name], dtype=dt)
            if len(col) != n:
                raise ValueError(f"column {name}: {len(col)} elements, runs add up to {n}")
            f.write(col.tobytes())
            f.write(b"\0" * (_pad(f.tell()) - f.tell()))


def read(path, mmap=True):
    """-> dict(kind, nsites, runs=[(name, count)], offsets=uint64[nruns+1], columns={name: array})."""
    with open(path, "rb") as f:
        head = f.read(_HEADER.size)
        if len(head) < _HEADER.size:
            raise ValueError(f"{path}: not a .pgtc file")
        magic, ver, kid, n, nruns, ncols, nbytes, data_off, _, _ = _HEADER.unpack(head)
        if magic != MAGIC or ver != 1 or kid not in _BY_ID:
            raise ValueError(f"{path}: not a .pgtc file (or unsupported version)")
        kind, spec = _BY_ID[kid]
        size = os.fstat(f.fileno()).st_size
        # same hostile-header checks as pgt_colfile.h open_view: no field may exceed the file before it is multiplied
        if nbytes > size or nruns > size // 8 or n > size or data_off > size or data_off % 4096:
            raise ValueError(f"{path}: corrupt header")
        counts = np.frombuffer(f.read(8 * nruns), dtype="<u8")
        names = f.read(nbytes).split(b"\0")[:nruns]
    if len(spec) != ncols or len(counts) != nruns or sum(int(c) for c in counts) != n or (nruns and int(counts.min()) == 0):
        raise ValueError(f"{path}: corrupt header")
    cols, at = {}, data_off
    for name, dt in spec:
        if at > size or n > (size - at) // np.dtype(dt).itemsize:
            raise ValueError(f"{path}: truncated file")
        if mmap and n:
            cols[name] = np.memmap(path, dtype=dt, mode="r", offset=at, shape=(n,))
        else:
            cols[name] = np.fromfile(path, dtype=dt, count=n, offset=at)
        at = _pad(at + n * np.dtype(dt).itemsize)
    return dict(kind=kind, nsites=n, runs=[(nm.decode(), int(c)) for nm, c in zip(names, counts)],
                offsets=np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64), columns=cols)


def _main(argv):
    """python -m popgenomicstools_b200.colfile info FILE.pgtc | head FILE.pgtc [N]"""
    if len(argv) < 2 or argv[0] not in ("info", "head"):
        print(_main.__doc__)
        return 2
    r = read(argv[1])
    if argv[0] == "info":
        print(f"kind\t{r['kind']}\nsites\t{r['nsites']}\nruns\t{len(r['runs'])}")
        print("columns\t" + ", ".join(f"{k}:{v.dtype}" for k, v in r["columns"].items()))
        for nm, c in r["runs"][:50]:
            print(f"run\t{nm}\t{c}")
        if len(r["runs"]) > 50:
            print(f"... {len(r['runs']) - 50} more runs")
        return 0
    n = int(argv[2]) if len(argv) > 2 else 10
    names = np.repeat(np.arange(len(r["runs"])), [c for _, c in r["runs"]])[:n]
    for i in range(min(n, r["nsites"])):
        print("\t".join([r["runs"][names[i]][0]] + [repr(v[i].item()) for v in r["columns"].values()]))
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(_main(sys.argv[1:]))
