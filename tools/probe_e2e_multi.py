"""Where does the one-process multi-GPU e2e lose H2D bandwidth?  (round 1: 2 processes 109 GB/s, round 2: one process
over 2 GPUs 76 GB/s.)   usage: probe_e2e_multi.py SITES DEVICES [raw]     e.g.  1e9 0,1   |   5e8 1   |   1e9 0,1 raw
Default: pgt_scan_sharded (C4 shape) from exactly-sized pinned columns, three timed calls.
raw: no library -- one host thread per device, torch non_blocking copies of 32 MB pieces from the same pinned columns."""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench as B  # noqa: E402
import popgenomicstools_b200 as pgt  # noqa: E402
from popgenomicstools_b200 import _cabi  # noqa: E402
from popgenomicstools_b200.workloads import human_like_contigs  # noqa: E402

n = int(float(sys.argv[1]))
devs = [int(x) for x in sys.argv[2].split(",")]
raw = len(sys.argv) > 3 and sys.argv[3] == "raw"
torch.cuda.set_device(devs[0])
names, offs = human_like_contigs(n, 10000)
t0 = time.perf_counter()
pinned = B.PinnedColumns(pgt, torch, "fst", 4, n, offs, torch.device(f"cuda:{devs[0]}"))
prep = time.perf_counter() - t0
if raw:
    a, b = torch.from_numpy(pinned.cols["a"]), torch.from_numpy(pinned.cols["b"])
    piece = 4 << 20  # elements = 32 MB
    per = n // len(devs)

    def work(k, d, out):
        torch.cuda.set_device(d)
        dst = torch.empty(2, piece, dtype=torch.float64, device=f"cuda:{d}")
        st = torch.cuda.Stream(device=d)
        lo, hi = k * per, (k + 1) * per
        with torch.cuda.stream(st):
            t = time.perf_counter()
            for s in range(lo, hi - piece, piece):
                dst[0].copy_(a[s:s + piece], non_blocking=True)
                dst[1].copy_(b[s:s + piece], non_blocking=True)
            st.synchronize()
            out[k] = time.perf_counter() - t

    for rep in range(2):
        out = [0.0] * len(devs)
        th = [threading.Thread(target=work, args=(k, d, out)) for k, d in enumerate(devs)]
        t = time.perf_counter()
        [x.start() for x in th]
        [x.join() for x in th]
        wall = time.perf_counter() - t
    print(f"raw devices={devs} n={n:.3g} wall={wall:.3f}s aggregate={16 * n / wall / 1e9:.1f} GB/s (torch pinned? {a.is_pinned()})", flush=True)
else:
    plan = pgt.WindowPlan(offs, 50000, 10000, unit_sites=512)
    out = pgt.scan_sharded(plan, _cabi.PGT_STAT_FST, pinned.cols, devs)
    ts = []
    for _ in range(3):
        t = time.perf_counter()
        pgt.scan_sharded(plan, _cabi.PGT_STAT_FST, pinned.cols, devs, out=out)
        ts.append(time.perf_counter() - t)
    best = min(ts)
    print(f"sharded devices={devs} n={n:.3g} prep={prep:.1f}s best={best:.3f}s aggregate={16 * n / best / 1e9:.1f} GB/s  all={[round(x, 3) for x in ts]}", flush=True)
pinned.free()
