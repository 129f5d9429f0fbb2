"""Small run of every extreme-scan kernel path (k_xunits G = 4..32, k_xwindows thread / warp combine,
device and host mode, shards, resident and uploaded tables) for compute-sanitizer memcheck."""
import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import popgenomicstools_b200 as pgt

rng = np.random.default_rng(0)
lengths = [30000, 12345, 7, 4001]
off = np.concatenate([[0], np.cumsum(lengths)]).astype(np.uint64)
n = int(off[-1])
pos = np.concatenate([np.cumsum(rng.integers(1, 9, size=L)) for L in lengths]).astype(np.uint32)
val = rng.normal(size=n)
val[::97] = np.nan
# unaligned device views
S = torch.empty(n + 1, dtype=torch.float64, device="cuda"); S[1:].copy_(torch.from_numpy(val))
P = torch.empty(n + 3, dtype=torch.int32, device="cuda"); P[3:].copy_(torch.from_numpy(pos.view(np.int32)))
for W in (1, 7, 100, 3000, 100000, 4000000000):
    for U in (0, 1, 5, 64):
        plan = pgt.ExtremePlan(pos, off, W, unit_sites=U)
        for g in (0, 4, 8, 16, 32):
            pgt.tune("xgroup", g)
            pgt.ihs_window(plan, P[3:], S[1:], 2.0)
            pgt.xpehh_window(plan, P[3:], S[1:], -1.0)
        pgt.tune("xgroup", 0)
        pgt.xpehh_window(plan, pos, val, 1.0)  # host mode
        for r in range(3):
            wl, wh, sl, sh = plan.shard(r, 3)
            if wh > wl:
                pgt.ihs_window(plan, P[3 + sl:], S[1 + sl:], 2.0, window_range=(wl, wh), site_origin=sl)
                pgt.ihs_window(plan, pos[sl:], val[sl:], 2.0, window_range=(wl, wh), site_origin=sl)
torch.cuda.synchronize()
print("sanitize extreme done")
