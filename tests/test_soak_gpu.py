"""Randomised differential soak of the scan against the oracle: random contig layouts (few long, many
short, lengths that trigger the cross-contig carry), random (W, S, unit), every statistic, device /
host / sharded, random kernel knobs.  Default size keeps it to a few seconds on a B200;
PGT_SOAK_CASES / PGT_SOAK_SEED scale it for one-off runs."""
import os

import numpy as np
import pytest

import oracle_lib as O
import parity as P
import textfmt as T

pytestmark = pytest.mark.gpu

CASES = int(os.environ.get("PGT_SOAK_CASES", "60"))
SEED = int(os.environ.get("PGT_SOAK_SEED", "0"))


def offsets(l):
    return np.concatenate([[0], np.cumsum(l)]).astype(np.uint64)


def npy(out):
    import torch
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}


def random_layout(rng, W, S):
    kind = rng.integers(0, 4)
    if kind == 0:    # a few long contigs
        lens = rng.integers(1, 60000, size=int(rng.integers(1, 5)))
    elif kind == 1:  # many short ones (more than 32 segments: tile -> segment table)
        lens = rng.integers(1, 1500, size=int(rng.integers(40, 300)))
    elif kind == 2:  # lengths around the window size, carry quirk (N - W) % S == 0
        lens = np.array([W + S * int(rng.integers(0, 5)) for _ in range(int(rng.integers(1, 8)))] + [int(rng.integers(1, W + 2))])
    else:            # mixed
        lens = np.concatenate([rng.integers(1, 40, size=10), rng.integers(1000, 30000, size=3), [W, W + S, max(1, W - S)]])
        rng.shuffle(lens)
    return [int(x) for x in lens]


def test_random_shapes_against_the_oracle():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import popgenomicstools_b200 as pgt
    rng = np.random.default_rng(20261018 + SEED)
    try:
        for it in range(CASES):
            W = int(rng.choice([1, 2, 7, 31, 64, 100, 255, 256, 257, 1000, 1048, 5000, 20000]))
            S = int(rng.choice([1, max(1, W // 10), max(1, W // 3), max(1, W - 1), W]))
            unit = int(rng.choice([0, 0, 32, 64, 512, 4096]))
            lengths = random_layout(rng, W, S)
            offs = offsets(lengths)
            n = int(offs[-1])
            if n // S > 400000:  # keep the oracle and the outputs small
                S = W
            chr_id = T.expand_chr(lengths)
            seed = int(rng.integers(1, 1000))
            density = int(rng.choice([1, 1, 3]))
            a, b = pgt.synth_fst(seed, 0, n)
            g = pgt.synth_het(seed, 0, n)
            f1, f2, n1, n2 = pgt.synth_dxy(seed, 0, n)
            pos = pgt.synth_pos(seed, 0, n, offs, density)
            h = {k: v.cpu().numpy() for k, v in dict(pos=pos, a=a, b=b, g=g, f1=f1, f2=f2, n1=n1, n2=n2).items()}
            minind = int(rng.integers(1, 12))
            rf = O.fst(chr_id, h["pos"], h["a"], h["b"], W, S)
            ra = O.fst(chr_id, h["pos"], np.abs(h["a"]), np.abs(h["b"]), W, S)
            rh = O.het(chr_id, h["pos"], h["g"], W, S)
            rd = O.dxy(chr_id, h["pos"], h["f1"], h["f2"], h["n1"], h["n2"], minind, W, S, 1)
            plan = pgt.WindowPlan(offs, W, S, unit_sites=unit)
            tag = f"case {it}: W={W} S={S} u={unit} contigs={len(lengths)} n={n}"
            assert plan.num_windows == len(rf["n"]), tag
            pgt.tune("level1", int(rng.choice([0, 0, 1, 2])))
            pgt.tune("level2", int(rng.choice([0, 0, 1, 2])))
            pgt.tune("slide", int(rng.choice([0, 0, 1, 2, 2])))  # 2: the sliding tile for every W <= 1048
            plan = pgt.WindowPlan(offs, W, S, unit_sites=unit)
            mode = rng.choice(["device", "host", "shards"])
            if mode == "device":
                res = npy(pgt.fused_window(plan, pos, a, b, g, f1, f2, n1, n2, minind=minind))
            elif mode == "host":
                res = pgt.fused_window(plan, h["pos"], h["a"], h["b"], h["g"], h["f1"], h["f2"], h["n1"], h["n2"], minind=minind)
            else:
                nsh = int(rng.integers(2, 6))
                parts = []
                for r in range(nsh):
                    wl, wh, sl, sh = plan.shard(r, nsh)
                    if wh > wl:
                        parts.append(npy(pgt.fused_window(plan, pos[sl:sh], a[sl:sh], b[sl:sh], g[sl:sh], f1[sl:sh], f2[sl:sh],
                                                          n1[sl:sh], n2[sl:sh], minind=minind, window_range=(wl, wh), site_origin=sl)))
                res = {k: np.concatenate([p[k] for p in parts]) for k in parts[0] if k != "dxy_global"} if parts else None
            if res is None or plan.num_windows == 0:
                continue
            tag += f" {mode}"
            P.assert_exact(res["label"], rf["label"], tag + " label")
            P.assert_exact(res["start_pos"], rf["start"], tag + " start")
            P.assert_exact(res["end_pos"], rf["end"], tag + " end")
            P.assert_exact(res["mid_pos"], rf["mid"], tag + " mid")
            P.assert_exact(res["nsites"], rf["n"], tag + " nsites")
            P.assert_sum_close(res["sum_a"], rf["asum"], ra["asum"], tag + " sum_a")
            P.assert_sum_close(res["sum_b"], rf["bsum"], ra["bsum"], tag + " sum_b")
            P.assert_exact(res["nhet"], rh["nhet"], tag + " nhet")
            P.assert_exact(res["nonmissing"], rh["nonmissing"], tag + " nonmissing")
            P.assert_exact(res["het"], rh["h"], tag + " het")
            P.assert_exact(res["neffective"], rd["neff"], tag + " neff")
            P.assert_exact(res["nskip"], rd["nskip"], tag + " nskip")
            P.assert_sum_close(res["dxy"], rd["dxy"], rd["dxy"], tag + " dxy")
    finally:
        pgt.tune("level1", 0)
        pgt.tune("level2", 0)
        pgt.tune("slide", 0)


def test_bp_mode_clustered_positions_against_the_oracle():
    """dxyWindow bp mode with very uneven SNP density: sparse chromosomes with dense stretches.  Tiles are
    sized for the AVERAGE density, so a dense stretch overflows a tile's shared-memory capacity and is read
    by the consumers straight from global memory (the `staged == false` path of k_units_tiled)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import popgenomicstools_b200 as pgt
    rng = np.random.default_rng(99 + SEED)
    try:
        for it in range(max(6, CASES // 6)):
            nchr = int(rng.integers(1, 4))
            chr_len, nsites, pos = [], [], []
            for _ in range(nchr):
                L = int(rng.integers(200_000, 1_500_000))
                p = [rng.choice(np.arange(1, L + 1), size=int(rng.integers(50, 2000)), replace=False)]  # sparse background
                for _ in range(int(rng.integers(1, 4))):  # dense stretches: every bp (or every 2nd) a site
                    st = int(rng.integers(1, L - 30000))
                    p.append(np.arange(st, st + int(rng.integers(3000, 25000)), int(rng.integers(1, 3))))
                p = np.unique(np.concatenate(p))
                p = p[p <= L]
                chr_len.append(L)
                nsites.append(len(p))
                pos.extend(p.tolist())
            soff = offsets(nsites)
            ns = int(soff[-1])
            W = int(rng.choice([100, 1000, 5000, 20000]))
            S = int(rng.choice([max(1, W // 10), max(1, W // 4), W]))
            unit = int(rng.choice([0, 0, 1024, 4096]))
            seed = int(rng.integers(1, 1000))
            f1, f2, n1, n2 = pgt.synth_dxy(seed, 0, ns)
            pd = torch.from_numpy(np.asarray(pos, np.int64).astype(np.int32)).cuda().view(torch.uint32)
            h = [np.asarray(pos, np.uint32)] + [x.cpu().numpy() for x in (f1, f2, n1, n2)]
            ref = O.dxy(T.expand_chr(nsites), h[0], h[1], h[2], h[3], h[4], 5, W, S, 0, 0, chr_len)
            plan = pgt.WindowPlan(offsets(chr_len), W, S, mode="bp", unit_sites=unit)
            tag = f"case {it}: W={W} S={S} u={unit} chr_len={chr_len} nsites={nsites}"
            assert plan.num_windows == len(ref["neff"]), tag
            for l1 in (0, 2):
                pgt.tune("level1", l1)
                res = npy(pgt.dxy_window(plan, pd, f1, f2, n1, n2, minind=5, site_offsets=soff))
                P.assert_exact(res["label"], ref["label"], tag)
                P.assert_exact(res["start_pos"].astype(np.int64), ref["start"], tag)
                P.assert_exact(res["end_pos"].astype(np.int64), ref["end"], tag)
                P.assert_exact(res["neffective"], ref["neff"], tag + f" neff l1={l1}")
                P.assert_exact(res["nskip"], ref["nskip"], tag + f" nskip l1={l1}")
                P.assert_sum_close(res["dxy"], ref["dxy"], ref["dxy"], tag + f" dxy l1={l1}")
            pgt.tune("level1", 0)
            hres = pgt.dxy_window(plan, *h, minind=5, site_offsets=soff)
            P.assert_exact(hres["neffective"], ref["neff"], tag + " host neff")
            P.assert_sum_close(hres["dxy"], ref["dxy"], ref["dxy"], tag + " host dxy")
    finally:
        pgt.tune("level1", 0)
