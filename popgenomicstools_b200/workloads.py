"""Synthetic workloads of BASELINE.json (shapes only; values come from the counter-based
generator, include/pgt_synth.h)."""
import numpy as np

# approximate GRCh38 chromosome sizes in Mb (chr1..22, X, Y): "human-like unequal lengths"
_HUMAN_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80, 59, 64, 47, 51,
             156, 57]


def human_like_contigs(n_total, stepsize):
    """24 contigs with human-like proportions summing to n_total sites.  Contig 1 is rounded to a
    multiple of `stepsize`, so (N - W) % S == 0 there and the reference's cross-contig carry
     is exercised at full scale.  Returns (names, offsets[25])."""
    tot = sum(_HUMAN_MB)
    lens = [int(n_total) * m // tot for m in _HUMAN_MB]
    lens[0] = max(stepsize, lens[0] // stepsize * stepsize)
    lens[-1] += int(n_total) - sum(lens)
    if lens[-1] <= 0:
        raise ValueError("n_total too small for 24 contigs")
    names = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]
    return names, np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)


WORKLOADS = {
    # BASELINE.json configs[3]: the configuration the metric is quoted on
    "C4": dict(stat="fst", n_sites=3_000_000_000, winsize=50000, stepsize=10000, seed=4, unit_sites=512,
               desc="fstWindow, 3e9 synthetic sites over 24 contigs, 50000-site windows / 10000-site step"),
    # BASELINE.json configs[4]: fused multi-stat scan, fine windows
    "C5": dict(stat="fused", n_sites=3_000_000_000, winsize=1000, stepsize=100, seed=5,
               desc="fused fst+dxy+het, 3e9 synthetic sites over 24 contigs, 1000-site windows / 100-site step"),
}
