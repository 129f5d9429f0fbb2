"""Shared parity helpers: tolerance of SURVEY.md §8(d).

Exact: window membership (first/last site), labels, positions, counts.
Sums:  |gpu - ref| <= 1e-9*|ref| + 1e-12*sum|x_i|   (the CUDA path adds in a fixed tree order,
       lane-strided partials + butterfly per unit, then units per window; the reference adds
       strictly left to right; see DESIGN.md "Summation order").
"""
import numpy as np

RTOL = 1e-9
ABS_COND = 1e-12


def assert_sum_close(gpu, ref, abs_sum, what):
    gpu, ref, abs_sum = np.asarray(gpu), np.asarray(ref), np.asarray(abs_sum)
    err = np.abs(gpu - ref)
    tol = RTOL * np.abs(ref) + ABS_COND * abs_sum
    bad = np.nonzero(~(err <= tol))[0]
    assert len(bad) == 0, f"{what}: {len(bad)} windows out of tolerance, first {bad[:5]}: gpu={gpu[bad[:5]]} ref={ref[bad[:5]]}"


def assert_exact(gpu, ref, what):
    gpu, ref = np.asarray(gpu), np.asarray(ref)
    assert gpu.shape == ref.shape, f"{what}: shape {gpu.shape} vs {ref.shape}"
    bad = np.nonzero(gpu != ref)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} mismatches, first at {bad[:5]}: gpu={gpu[bad[:5]]} ref={ref[bad[:5]]}"


def rows_match_modulo_ties(rows_gpu, rows_ref, float_cols):
    """Text rows equal except last-digit %g rounding ties in the float columns; returns #ties."""
    assert len(rows_gpu) == len(rows_ref), (len(rows_gpu), len(rows_ref))
    ties = 0
    for rg, rr in zip(rows_gpu, rows_ref):
        if rg == rr:
            continue
        fg, fr = rg.split("\t"), rr.split("\t")
        assert len(fg) == len(fr), (rg, rr)
        for i, (x, y) in enumerate(zip(fg, fr)):
            if x == y:
                continue
            assert i in float_cols, (rg, rr)
            xv, yv = float(x), float(y)
            assert abs(xv - yv) <= 2e-6 * max(abs(xv), abs(yv)), (rg, rr)
            ties += 1
    return ties
