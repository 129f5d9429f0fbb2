"""Scratch probe: level-1/level-2 timing of the fst scan at several sizes (not the bench)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import popgenomicstools_b200 as pgt

def contigs(n_total, S):
    mb = [248,242,198,190,182,171,159,145,138,134,135,133,114,107,102,90,83,80,59,64,47,51,156,57]
    tot = sum(mb); L = [n_total*m//tot for m in mb]
    L[0] = max(S, L[0]//S*S)
    L[-1] += n_total - sum(L)
    return np.concatenate([[0], np.cumsum(L)]).astype(np.uint64)

for n, W, S, unit in [(int(float(x.split(',')[0])), int(x.split(',')[1]), int(x.split(',')[2]), int(x.split(',')[3])) for x in sys.argv[1:]]:
    offs = contigs(n, S)
    plan = pgt.WindowPlan(offs, W, S, unit_sites=unit)
    a, b = pgt.synth_fst(4, 0, n)
    pos = pgt.synth_pos(4, 0, n, offs, 1)
    torch.cuda.synchronize()
    out = pgt.fst_window(plan, pos, a, b)
    for _ in range(3): pgt.fst_window(plan, pos, a, b, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    K = 10
    ev[0].record()
    for _ in range(K): pgt.fst_window(plan, pos, a, b, out=out)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / K
    print(f"n={n:.3g} W={W} S={S} u={unit} windows={plan.num_windows} units={plan.num_units} ms={ms:.4f} "
          f"sites/s={n/ms*1e3:.4g} GB/s(16B)={16*n/ms/1e6:.1f} GB/s(20B)={20*n/ms/1e6:.1f}", flush=True)
    del a, b, pos, out, plan
    torch.cuda.empty_cache()
