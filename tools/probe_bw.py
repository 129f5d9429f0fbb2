"""Scratch probe: level-1/level-2 timing at several (stat, n, W, S, unit, level1) points (not the bench).
usage: probe_bw.py stat,n,W,S,unit,level1[,stages,stage_kb[,level2]] ...
stat "dxybpD" = dxyWindow bp mode (-fixedsite 0) with one site per D bp, W and S in bp."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import popgenomicstools_b200 as pgt
from popgenomicstools_b200 import _cabi
from popgenomicstools_b200.workloads import human_like_contigs

BPS = {"fst": 16, "het": 1, "dxy": 24, "fused": 41, "dxybp": 28}
STAT = {"fst": _cabi.PGT_STAT_FST, "het": _cabi.PGT_STAT_HET, "dxy": _cabi.PGT_STAT_DXY, "fused": _cabi.PGT_STAT_FUSED,
        "dxybp": _cabi.PGT_STAT_DXY}
for spec in sys.argv[1:]:
    f = spec.split(",")
    stat, n, W, S, unit, l1 = f[:6]
    n, W, S, unit, l1 = int(float(n)), int(W), int(S), int(unit), int(l1)
    stages, skb = (int(f[6]), int(f[7])) if len(f) > 7 else (2, 110)
    l2 = int(f[8]) if len(f) > 8 else 0
    pgt.tune("level2", l2)
    pgt.tune("stages", stages); pgt.tune("stage_kb", skb)
    _, offs = human_like_contigs(n, S)
    import os
    if os.environ.get("PROBE_CONTIGS"):  # many equal contigs (scaffold-level assemblies)
        nc = int(os.environ["PROBE_CONTIGS"])
        lens = np.full(nc, n // nc, np.int64) + (np.arange(nc) % 7)  # ragged by a few sites
        n = int(lens.sum())
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    extra = {}
    if stat.startswith("dxybp"):
        dens = int(stat[5:] or 1)
        stat = "dxybp"
        lens = np.diff(offs).astype(np.int64) * dens + 17  # chromosome lengths in bp
        plan = pgt.WindowPlan(np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64), W, S, mode="bp", unit_sites=unit)
        cols = dict(pos=pgt.synth_pos(4, 0, n, offs, dens))
        extra = dict(site_offsets=offs)
    else:
        plan = pgt.WindowPlan(offs, W, S, unit_sites=unit)
        cols = dict(pos=pgt.synth_pos(4, 0, n, offs, 1))
    if stat in ("fst", "fused"):
        cols["a"], cols["b"] = pgt.synth_fst(4, 0, n)
    if stat in ("het", "fused"):
        cols["geno"] = pgt.synth_het(4, 0, n)
    if stat in ("dxy", "fused", "dxybp"):
        cols["f1"], cols["f2"], cols["n1"], cols["n2"] = pgt.synth_dxy(4, 0, n)
    pgt.tune("level1", l1)
    torch.cuda.synchronize()
    out = pgt.scan(plan, STAT[stat], cols, minind=5, **extra)
    for _ in range(3):
        pgt.scan(plan, STAT[stat], cols, minind=5, out=out, **extra)
    torch.cuda.synchronize()
    pgt.profile(True); pgt.profile_read()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    K = 10
    ev[0].record()
    for _ in range(K):
        pgt.scan(plan, STAT[stat], cols, minind=5, out=out, **extra)
    ev[1].record(); torch.cuda.synchronize()
    pr = pgt.profile_read(); pgt.profile(False)
    ms = ev[0].elapsed_time(ev[1]) / K
    l1ms = pr["units_ms"] / K; l2ms = pr["windows_ms"] / K
    print(f"{stat} n={n:.3g} W={W} S={S} u={unit} l1={l1} l2={l2} st={stages}x{skb}KB win={plan.num_windows} units={plan.num_units} step_ms={ms:.4f} "
          f"L1_ms={l1ms:.4f} L2_ms={l2ms:.4f} sites/s={n/ms*1e3:.4g} L1_GB/s={BPS[stat]*n/max(l1ms,1e-9)/1e6:.1f}", flush=True)
    del cols, out, plan
    torch.cuda.empty_cache()
