#!/usr/bin/env python3
"""Secondary bench: the ihsWindow / xpehhWindow extreme-score scan (pgt_scan_extreme) on one B200.

    python tools/bench_extreme.py [--sites 1e9] [--density 3] [--winsize 100000,1000,100] [--steps 20]

Prints one JSON line per window size: whole-scan sites/s (CUDA events, scores resident in HBM),
the level-1 kernel's roofline (8 B/site algorithmic vs MEASURED_PEAKS.json), the host bookkeeping
time (pgt_xplan_create), e2e with host columns, and the CPU oracle / reference binary next to it.
Not the BASELINE metric (that is bench.py); evidence for SURVEY.md section 8f rank 3.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=float, default=1e9)
    ap.add_argument("--density", type=int, default=3)
    ap.add_argument("--winsize", default="100000,1000,100")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cpu-sites", type=float, default=2e6)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--xgroup", default="0", help="comma list of forced lanes-per-unit values (0 = auto)")
    a = ap.parse_args()
    import numpy as np
    import torch
    import popgenomicstools_b200 as pgt
    from popgenomicstools_b200.workloads import human_like_contigs
    import bench
    n = int(a.sites)
    names, offs = human_like_contigs(n, 1)
    dev = torch.device("cuda:0")
    pos_d = pgt.synth_pos(6, 0, n, offs, a.density, device=dev)
    score_d = pgt.synth_score(6, 0, n, device=dev)
    torch.cuda.synchronize()
    pos_h = pos_d.cpu().numpy()
    peak, peak_src = bench.hbm_peak()
    for W, xg in [(int(x), int(g)) for x in a.winsize.split(",") for g in a.xgroup.split(",")]:
        pgt.tune("xgroup", xg)
        t0 = time.perf_counter()
        plan = pgt.ExtremePlan(pos_h, offs, W)
        plan_s = time.perf_counter() - t0
        out = pgt.ihs_window(plan, pos_d, score_d, 2.0)
        for _ in range(a.warmup):
            pgt.ihs_window(plan, pos_d, score_d, 2.0, out=out)
        torch.cuda.synchronize()
        pgt.profile(True)
        pgt.profile_read_extreme()
        l0 = pgt.kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            pgt.ihs_window(plan, pos_d, score_d, 2.0, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        prof = pgt.profile_read_extreme()
        pgt.profile(False)
        l1_ms = prof["units_ms"] / max(1, prof["units_launches"])
        l2_ms = prof["windows_ms"] / max(1, prof["windows_launches"])
        algo = 8 * n
        line = {"metric": "sites/sec of the ihsWindow extreme-score window scan on 1 B200", "value": n / (ms * 1e-3),
                "unit": "sites/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms, 4),
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"ihsWindow, {n} synthetic sites over 24 chromosomes (1 SNP per {a.density} bp), {W}-bp windows, cutoff 2",
                           "windows": plan.num_windows, "units": plan.num_units, "unit_sites": 2048, "xgroup": xg,
                           "l2": "8 GB score column >> 126 MB L2; no flush needed"},
                "gpu_launches": pgt.kernel_launch_count() - l0,
                "host_bookkeeping_s": round(plan_s, 3),
                "roofline": {"bound": "hbm", "achieved": round(algo / (l1_ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(algo / (l1_ms * 1e-3) / 1e9 / peak, 4), "traffic": None, "kernel": "k_xunits (level 1)",
                             "peak_source": peak_src, "algorithmic_bytes_per_launch": algo, "bytes_per_site": 8,
                             "kernel_ms_per_launch": round(l1_ms, 4), "level2_ms_per_launch": round(l2_ms, 4)}}
        # e2e: host columns through the C ABI (H2D of the score column + D2H of the rows inside the timed region)
        try:
            if a.no_e2e:
                raise RuntimeError('skipped (--no-e2e)')
            hs = torch.empty(n, dtype=torch.float64, pin_memory=True)
            hs.copy_(score_d)
            torch.cuda.synchronize()
            hout = pgt.ihs_window(plan, pos_h, hs.numpy(), 2.0)
            for k in ("ext_value", "ext_pos", "nbig", "nsites"):
                assert hout[k].tobytes() == out[k].cpu().numpy().tobytes(), k
            t0 = time.perf_counter()
            for _ in range(3):
                pgt.ihs_window(plan, pos_h, hs.numpy(), 2.0, out=hout)
            e_ms = (time.perf_counter() - t0) / 3 * 1e3
            line["e2e"] = {"value": n / (e_ms * 1e-3), "unit": "sites/s", "h2d_bytes_per_step": 8 * n,
                           "d2h_bytes_per_step": int(sum(v.nbytes for k, v in hout.items() if k != "ext_pos")),
                           "ms_per_step": round(e_ms, 3), "h2d_gbs": round(8 * n / (e_ms * 1e-3) / 1e9, 2)}
            del hs
        except Exception as ex:
            line["e2e"] = {"value": None, "error": repr(ex)[:200]}
        if not a.no_cpu:
            import oracle_lib as O
            import textfmt as T
            m = int(min(a.cpu_sites, n))
            cn, coffs = human_like_contigs(m, 1)
            cpos = O.synth_pos(6, coffs, a.density)
            cval = O.synth_score(6, 0, m)
            chr_id = np.repeat(np.arange(24, dtype=np.uint32), np.diff(coffs).astype(np.int64))
            t0 = time.perf_counter()
            O.extreme("ihs", chr_id, cpos, cval, W, 2.0)
            osec = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": round(m / osec, 1), "unit": "sites/s", "cores": 1, "kind": "port",
                                    "sample": f"oracle/pgt_oracle_extreme.c compute-only on {m} pre-parsed sites"}
            ref = O.ref_binary("ihsWindow")
            if ref:
                with tempfile.TemporaryDirectory() as d:
                    p = os.path.join(d, "s.norm")
                    lens = np.diff(coffs).astype(np.int64).tolist()
                    open(p, "w").write(T.ihs_text(cn, lens, cpos, np.round(cval * 1e6).astype(np.int64)))
                    t0 = time.perf_counter()
                    r = subprocess.run([ref, p, "-winsize", str(W)], capture_output=True, text=True)
                    rsec = time.perf_counter() - t0
                    ours = os.path.join(ROOT, "popgenomicstools_b200", "bin", "ihsWindow")
                    best = None
                    for _ in range(2):
                        t0 = time.perf_counter()
                        g = subprocess.run([ours, p, "-winsize", str(W)], capture_output=True, text=True,
                                           env=dict(os.environ, PGT_TIMING="1"))
                        best = time.perf_counter() - t0
                    line["cpu_baseline"]["reference_binary"] = {
                        "sites": m, "wall_s": round(rsec, 3), "sites_per_s": round(m / rsec, 1),
                        "our_cli_wall_s": round(best, 3), "our_cli_timing": json.loads(g.stderr.strip().splitlines()[-1]),
                        "identical_stdout": g.stdout == r.stdout, "rows": len(r.stdout.splitlines())}
        print(json.dumps(line), flush=True)
        del plan, out


if __name__ == "__main__":
    main()
