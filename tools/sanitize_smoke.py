"""Small run of every kernel path for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import popgenomicstools_b200 as pgt
import oracle_lib as O

def offsets(l): return np.concatenate([[0], np.cumsum(l)]).astype(np.uint64)

lengths = [5000 + 300, 12345, 7, 4001]
offs = offsets(lengths); n = int(offs[-1])
a, b = pgt.synth_fst(1, 0, n); g = pgt.synth_het(1, 0, n); f1, f2, n1, n2 = pgt.synth_dxy(1, 0, n)
pos = pgt.synth_pos(1, 0, n, offs, 1)
# unaligned views
A = torch.empty(n + 3, dtype=torch.float64, device="cuda"); A[3:].copy_(a)
G = torch.empty(n + 5, dtype=torch.int8, device="cuda"); G[5:].copy_(g)
for (W, S, u) in [(5000, 100, 0), (1000, 100, 0), (300, 299, 0), (1, 1, 0), (64, 1, 0), (777, 13, 32), (2560, 256, 0), (4096, 4096, 4096)]:
    plan = pgt.WindowPlan(offs, W, S, unit_sites=u)
    for l1 in (0, 1, 2):
        pgt.tune("level1", l1)
        for l2 in (0, 1, 2):
            pgt.tune("level2", l2)
            pgt.fst_window(plan, pos, A[3:], b)
            pgt.het_window(plan, pos, G[5:])
            pgt.dxy_window(plan, pos, f1, f2, n1, n2, minind=5)
            pgt.fused_window(plan, pos, A[3:], b, G[5:], f1, f2, n1, n2, minind=5)
    pgt.tune("level1", 0); pgt.tune("level2", 0)
    # shards on views
    for r in range(3):
        wl, wh, sl, sh = plan.shard(r, 3)
        if wh > wl:
            pgt.fst_window(plan, pos[sl:sh], a[sl:sh], b[sl:sh], window_range=(wl, wh), site_origin=sl)
    # host mode
    pgt.fst_window(plan, pos.cpu().numpy(), a.cpu().numpy(), b.cpu().numpy())
    pgt.fused_window(plan, pos.cpu().numpy(), a.cpu().numpy(), b.cpu().numpy(), g.cpu().numpy(), f1.cpu().numpy(), f2.cpu().numpy(),
                     n1.cpu().numpy(), n2.cpu().numpy(), minind=5)
torch.cuda.synchronize()
# bp mode, sparse + dense, device + host
nsites = [3000, 1500, 40]
soff = offsets(nsites); ns = int(soff[-1])
for density in (10, 1):
    chr_len = [x * density + 17 for x in nsites]
    f1, f2, n1, n2 = pgt.synth_dxy(2, 0, ns); p = pgt.synth_pos(2, 0, ns, soff, density)
    for (W, S) in [(2000, 500), (100, 100), (1, 1), (777, 10)]:
        plan = pgt.WindowPlan(offsets(chr_len), W, S, mode="bp")
        for l1 in (0, 2):
            pgt.tune("level1", l1)
            pgt.dxy_window(plan, p, f1, f2, n1, n2, minind=5, site_offsets=soff)
        pgt.tune("level1", 0)
        pgt.dxy_window(plan, p.cpu().numpy(), f1.cpu().numpy(), f2.cpu().numpy(), n1.cpu().numpy(), n2.cpu().numpy(), minind=5, site_offsets=soff)
torch.cuda.synchronize()
print("sanitize smoke done")
