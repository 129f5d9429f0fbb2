"""Scratch probe: host-memory mode (PGT_MEM_HOST) of the fst scan from PAGEABLE columns (what the CLIs pass:
malloc'ed arrays or the mapping of a .pgtc cache) against pinned columns (what bench.py's e2e passes).
Prints sites/s and the implied H2D GB/s for both; the gap is what a pinned staging ring inside the
library's host mode could recover for the CLIs at genome scale."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import popgenomicstools_b200 as pgt

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
offs = np.array([0, n], np.uint64)
plan = pgt.WindowPlan(offs, 50000, 10000, unit_sites=512)
a_d, b_d = pgt.synth_fst(4, 0, n, device="cuda")
pos_d = pgt.synth_pos(4, 0, n, offs, 1, device="cuda")
cols = {}
for kind in ("pinned", "pageable"):
    pin = kind == "pinned"
    a = torch.empty(n, dtype=torch.float64, pin_memory=pin); a.copy_(a_d)
    b = torch.empty(n, dtype=torch.float64, pin_memory=pin); b.copy_(b_d)
    p = torch.empty(n, dtype=torch.int32, pin_memory=pin); p.copy_(pos_d.view(torch.int32))
    cols[kind] = (p.numpy().view(np.uint32), a.numpy(), b.numpy())
torch.cuda.synchronize()
ref = None
for kind, stage in (("pinned", 0), ("pageable", 0), ("pageable", 1), ("pinned", 0), ("pageable", 0), ("pageable", 1)):
    pgt.tune("hoststage", stage)  # 0 = plain cudaMemcpyAsync from pageable memory (default), 1 = experimental pinned ring
    p, a, b = cols[kind]
    out = pgt.fst_window(plan, p, a, b)  # warm
    t0 = time.perf_counter()
    for _ in range(3):
        out = pgt.fst_window(plan, p, a, b)
    dt = (time.perf_counter() - t0) / 3
    s = np.asarray(out["sum_a"]).copy()
    if ref is None:
        ref = s
    tag = kind + (" (pinned ring)" if stage == 1 else " (driver staging)" if kind == "pageable" else "")
    print(f"{tag:26s}: {n / dt:.3e} sites/s, {16 * n / dt / 1e9:.1f} GB/s H2D, identical={np.array_equal(s, ref)}", flush=True)
