#!/usr/bin/env python3
"""bench.py -- sites/sec of the fused window scan on 1..8 B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4|C5] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step is one pass of the hot path (level-1 unit reduction + level-2 window combine, and for
N > 1 the NCCL gather of the per-window results to rank 0) over the whole synthetic genome,
sharded by site range with a (W-S)-site halo (strong scaling: total work fixed).  Inputs are
generated on the device by the counter-based generator and are resident in HBM when the timed
region starts; they are far larger than L2 (48 GB vs 126 MB), so no explicit flush is needed.

Prints ONE JSON line (rank 0).  `value` = total sites / max-over-ranks device time;
`e2e` = the same scan through the C ABI with HOST (pinned) columns, H2D/D2H inside the timed
region; `roofline` = algorithmic bytes of the dominant kernel (level 1) / its mean CUDA-event
duration, against MEASURED_PEAKS.json; `cpu_baseline` = the unmodified reference binary
(oracle/_ref, g++ -O3) timed on this box's host cores on a bounded text sample.

--impl reference runs only that CPU arm (the reference's own implementation of the path).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sites/sec for fused FST/dxy window scan at 1/2/4/8 B200; % of HBM peak"
UNIT = "sites/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=["C4", "C5"])
    ap.add_argument("--sites", type=float, default=0, help="override the workload's total site count (debug)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-sites", type=float, default=4.8e7)
    return ap.parse_args()


# ------------------------------------------------------------------------------ reference arm

def _write_contig(job):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import numpy as np
    path, name, lo, hi, seed = job
    O.write_text("fst", path, [name], np.array([lo, hi], np.uint64), seed=seed, density=1)
    # write_text restarts local positions at 1 for the contig and hashes by global site index
    return path


def _run_ref(job):
    exe, path, W, S = job
    t0 = time.perf_counter()
    subprocess.run([exe, path, str(W), str(S)], stdout=subprocess.DEVNULL, check=True)
    return time.perf_counter() - t0


def cli_arm(tmp, jobs, W, S):
    """Our drop-in fstWindow on the same text, concatenated into one whole-genome file as a user
    would run it: text parsing (timed separately), scan through the C ABI, row formatting."""
    exe = os.path.join(ROOT, "popgenomicstools_b200", "bin", "fstWindow")
    if not os.access(exe, os.X_OK):
        return None
    allp = os.path.join(tmp, "all.fst")
    with open(allp, "wb") as w:
        for j in jobs:
            with open(j[0], "rb") as r:
                shutil.copyfileobj(r, w, 1 << 24)
    best = None
    ours_out = os.path.join(tmp, "ours.tsv")
    for _ in range(2):  # second run = warm page cache + warm driver
        t0 = time.perf_counter()
        with open(ours_out, "wb") as fo:
            p = subprocess.run([exe, allp, str(W), str(S)], stdout=fo, stderr=subprocess.PIPE, text=True,
                               env=dict(os.environ, PGT_TIMING="1"))
        wall = time.perf_counter() - t0
        if p.returncode != 0:
            return {"error": p.stderr[-200:]}
        t = json.loads(p.stderr.strip().splitlines()[-1])
        t["wall_s_incl_process_and_cuda_startup"] = round(wall, 3)
        best = t
    best["sites_per_s_parse_scan_format"] = round(best["sites"] / (best["total_ms"] * 1e-3), 1)
    # the same input as a binary columnar cache (PGT_PACK, csrc/tools/pgt_colfile.h): no text parsing at all
    try:
        cache = os.path.join(tmp, "all.pgtc")
        t0 = time.perf_counter()
        subprocess.run([exe, allp, str(W), str(S)], check=True, env=dict(os.environ, PGT_PACK=cache))
        pack_s = time.perf_counter() - t0
        cache_out = os.path.join(tmp, "cache.tsv")
        ct = None
        for _ in range(2):
            t0 = time.perf_counter()
            with open(cache_out, "wb") as fo:
                p = subprocess.run([exe, cache, str(W), str(S)], stdout=fo, stderr=subprocess.PIPE, text=True,
                                   env=dict(os.environ, PGT_TIMING="1"), check=True)
            cwall = time.perf_counter() - t0
            ct = json.loads(p.stderr.strip().splitlines()[-1])
        best["columnar_cache"] = {"pack_s": round(pack_s, 3), "file_bytes": os.path.getsize(cache), "load_ms": ct["parse_ms"],
                                  "scan_ms": ct["scan_ms"], "format_ms": ct["format_ms"], "total_ms": ct["total_ms"],
                                  "wall_s_incl_process_and_cuda_startup": round(cwall, 3),
                                  "identical_stdout": open(cache_out, "rb").read() == open(ours_out, "rb").read()}
        os.remove(cache)
    except Exception as ex:
        best["columnar_cache"] = {"error": repr(ex)[:200]}
    # the same file through the unmodified reference binary, one process (how a user runs it), and a
    # row-by-row comparison of the two outputs (text after %g formatting)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    ref = O.ref_binary("fstWindow")
    if ref:
        ref_out = os.path.join(tmp, "ref.tsv")
        t0 = time.perf_counter()
        with open(ref_out, "wb") as fo:
            subprocess.run([ref, allp, str(W), str(S)], stdout=fo, check=True)
        rsec = time.perf_counter() - t0
        a, b = open(ours_out).read().splitlines(), open(ref_out).read().splitlines()
        diff = sum(1 for x, y in zip(a, b) if x != y) + abs(len(a) - len(b))
        best["reference_same_file_one_process"] = {"wall_s": round(rsec, 2), "sites_per_s": round(best["sites"] / rsec, 1),
                                                   "rows": len(b), "rows_differing_from_ours": diff}
    os.remove(allp)
    return best


def reference_arm(wl, sample_sites, steps, warmup, with_cli=False):
    """Times the unmodified reference fstWindow (oracle/_ref, built from /root/reference with
    g++ -O3 -Wall) end to end -- text parsing is inseparable from its window arithmetic -- on a
    scaled twin of the workload: the same 24 contig-length ratios, one text file and one
    single-threaded reference process per contig, min(24, cores) processes at a time."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from popgenomicstools_b200.workloads import human_like_contigs
    W, S = wl["winsize"], wl["stepsize"]
    exe = O.ref_binary("fstWindow")
    cores = os.cpu_count() or 1
    names, offs = human_like_contigs(int(sample_sites), S)
    n = int(offs[-1])
    if exe is None:
        # reference binary not built (oracle/_ref absent): time the oracle port, compute only
        import numpy as np
        a, b = O.synth_fst(wl["seed"], 0, n)
        pos = O.synth_pos(wl["seed"], offs, 1)
        chr_id = np.repeat(np.arange(24, dtype=np.uint32), np.diff(offs).astype(np.int64))
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            O.fst(chr_id, pos, a, b, W, S, count_only=True)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        sec = sum(times) / len(times)
        return dict(value=n / sec, unit=UNIT, cores=1, kind="port",
                    sample=f"oracle/pgt_oracle.c compute-only (pre-parsed arrays), {n} sites, 24 contigs, {W}/{S}"), sec
    tmp = tempfile.mkdtemp(prefix="pgt_ref_")
    try:
        O.build_oracle()
        jobs = [(os.path.join(tmp, f"{nm}.fst"), nm, int(offs[i]), int(offs[i + 1]), wl["seed"]) for i, nm in enumerate(names)]
        nproc = min(len(jobs), cores)
        with mp.get_context("fork").Pool(nproc) as pool:
            pool.map(_write_contig, jobs)
            rjobs = [(exe, j[0], W, S) for j in sorted(jobs, key=lambda j: j[2] - j[3])]  # longest first
            times = []
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                pool.map(_run_ref, rjobs, chunksize=1)
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
        sec = sum(times) / len(times)
        tb = sum(os.path.getsize(j[0]) for j in jobs)
        cli = cli_arm(tmp, jobs, W, S) if with_cli else None
        return dict(value=n / sec, unit=UNIT, cores=nproc, kind="reference", **({"our_cli_same_text": cli} if cli else {}),
                    sample=(f"unmodified reference fstWindow (g++ -O3 -Wall) end to end incl. text parsing: scaled twin "
                            f"{n} sites / 24 contig files ({tb / 1e9:.2f} GB text), {W}/{S}, one process per contig, "
                            f"{nproc} at a time on {cores} host cores")), sec
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from popgenomicstools_b200.workloads import WORKLOADS
    wl = WORKLOADS["C4"] if args.workload == "C5" else WORKLOADS[args.workload]  # the reference has no fused tool
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # bound the whole run to a few minutes whatever K/W the driver passes
    per_step_budget = 240.0 / (steps + warmup)
    sample = min(args.cpu_sample_sites, max(2.4e6, per_step_budget * 1.0e7))
    cb, sec = reference_arm(wl, sample, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["desc"], "sample": cb["sample"]},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------ clocks

class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.005)

    def __enter__(self):
        if self._h is not None:
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._h is not None:
            self._t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "window": "warm-up + timed steps (one contiguous load)"}


# ------------------------------------------------------------------------------ B200 arm

def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload, n_gpus):
    """dram bytes per level-1 launch from the committed ncu --set full capture, if it matches."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if n_gpus != 1 or not os.path.exists(p):
        return None
    try:
        return json.load(open(p)).get(workload, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import popgenomicstools_b200 as pgt
    from popgenomicstools_b200 import _cabi
    from popgenomicstools_b200.workloads import WORKLOADS, human_like_contigs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    wl = WORKLOADS[args.workload]
    n_total = int(args.sites) if args.sites else wl["n_sites"]
    W, S, seed = wl["winsize"], wl["stepsize"], wl["seed"]
    fused = wl["stat"] == "fused"
    names, offs = human_like_contigs(n_total, S)
    # reduction unit (part of the summation order, any value gives the same results within 1e-16):
    # 512 sites fills a 110 KB tile of the 16 B/site fst stream with 13 units = one round of the 15
    # consumer warps (+2 % over the default 256, measured); the 41 B/site fused tile is best at 256
    unit_sites = wl.get("unit_sites", 0)
    plan = pgt.WindowPlan(offs, W, S, unit_sites=unit_sites)
    w_lo, w_hi, s_lo, s_hi = plan.shard(rank, world)
    n_local, nwin_local = s_hi - s_lo, w_hi - w_lo

    # ---- resident synthetic columns of this rank's shard (incl. halo)
    pos = pgt.synth_pos(seed, s_lo, n_local, offs, 1, device=dev)
    a, b = pgt.synth_fst(seed, s_lo, n_local, device=dev)
    cols = dict(pos=pos, a=a, b=b)
    stat = _cabi.PGT_STAT_FST
    bytes_per_site = 16
    if fused:
        stat = _cabi.PGT_STAT_FUSED
        cols["geno"] = pgt.synth_het(seed, s_lo, n_local, device=dev)
        cols["f1"], cols["f2"], cols["n1"], cols["n2"] = pgt.synth_dxy(seed, s_lo, n_local, device=dev)
        bytes_per_site = 41
    torch.cuda.synchronize()

    # ---- outputs live in one packed buffer so the gather to rank 0 is a single NCCL call
    from popgenomicstools_b200.scan import _STAT_OUTS
    from popgenomicstools_b200.sharding import PackedWindows, shard_counts
    fields = [k for k in _STAT_OUTS[stat] if k != "dxy_global"]
    counts = shard_counts(dist, nwin_local, world, dev, torch)
    pw = PackedWindows(_STAT_OUTS[stat], nwin_local, max(counts), dev, torch)
    out = pw.views
    gathered = [None]

    def step():
        pgt.scan(plan, stat, cols, minind=5, window_range=(w_lo, w_hi), site_origin=s_lo, out=out)
        if world > 1:
            gathered[0] = pw.gather(dist, rank, world, gathered[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The clock sampler runs from the first warm-up step to the end of the timed region: NVML
    # queries take milliseconds, and at 8 GPUs the timed region itself is only ~20 ms.  Warm-up is
    # the same load, contiguous with the timed steps, and lasts at least 0.3 s.
    n_warm = 0
    with ClockSampler(local) as clk:
        t_w = time.perf_counter()
        for _ in range(max(3, args.warmup)):
            step()
        n_warm = max(3, args.warmup)
        barrier()
        per = torch.tensor([(time.perf_counter() - t_w) / n_warm], device=dev, dtype=torch.float64)
        if world > 1:  # every rank must issue the same number of collectives
            dist.all_reduce(per, op=dist.ReduceOp.MAX)
        extra = max(0, min(2000, int(0.3 / max(float(per.item()), 1e-5)) - n_warm))
        for _ in range(extra):
            step()
        n_warm += extra
        nw = torch.tensor([n_warm], device=dev)
        barrier()
        launches0 = pgt.kernel_launch_count()
        pgt.profile(True)
        pgt.profile_read()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
    ms_total = e0.elapsed_time(e1)
    prof = pgt.profile_read()
    pgt.profile(False)
    launches = pgt.kernel_launch_count() - launches0
    t = torch.tensor([ms_total, float(launches)], device=dev, dtype=torch.float64)
    tmax = t.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms_step = float(tmax[0].item()) / args.steps
    total_launches = int(t[1].item())
    value = n_total / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (level 1) on this rank
    peak, peak_src = hbm_peak()
    nunits_local = 0
    if nwin_local:
        fu0, _ = plan.window_units(w_lo)
        fu1, c1 = plan.window_units(w_hi - 1)
        nunits_local = (plan.num_units if w_hi == plan.num_windows else fu1 + c1) - (0 if w_lo == 0 else fu0)
    acc_bytes = 40 if fused else 16
    algo_bytes = bytes_per_site * n_local + acc_bytes * nunits_local
    l1_ms = prof["units_ms"] / max(1, prof["units_launches"])
    achieved = algo_bytes / (l1_ms * 1e-3) / 1e9 if l1_ms > 0 else 0.0
    ach = torch.tensor([achieved, prof["units_ms"], prof["windows_ms"]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ach, op=dist.ReduceOp.SUM)
        ach /= world
    roofline = {"bound": "hbm", "achieved": round(float(ach[0].item()), 1), "peak": peak, "unit": "GB/s",
                "frac": round(float(ach[0].item()) / peak, 4), "traffic": recorded_traffic(args.workload, world),
                "kernel": "k_units (level 1: per-site statistic + unit reduction)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo_bytes,
                "bytes_per_site": bytes_per_site,
                "kernel_ms_per_launch": round(l1_ms, 4),
                "kernel_share_of_step": round(float(ach[1].item()) / args.steps / ms_step, 4),
                "level2_ms_per_launch": round(float(ach[2].item()) / args.steps, 4),
                "note": ("pos is gathered only at the two edges of each window, so the compulsory stream is "
                         "a+b = 16 B/site (SURVEY.md 8d conservative variant)" if not fused else
                         "a,b,f1,f2 f64 + n1,n2 i32 + genotype i8 = 41 B/site; pos gathered at window edges only")}
    if not fused:
        roofline["achieved_if_pos_counted_20B"] = round(float(ach[0].item()) * 20 / 16, 1)

    # ---- end to end through the C ABI with host (pinned) columns
    e2e = None
    if not args.no_e2e:
        try:
            hcols = {}
            for k, v in cols.items():
                if k == "pos":
                    hcols[k] = v.cpu().numpy()  # gathered on the host at window edges, never copied to the device
                else:
                    h = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
                    h.copy_(v)
                    hcols[k] = h.numpy()
            torch.cuda.synchronize()
            hout = pgt.scan(plan, stat, hcols, minind=5, window_range=(w_lo, w_hi), site_origin=s_lo, device=dev)
            for k in fields:  # the host path returns exactly what the resident path computed
                assert hout[k].tobytes() == out[k].cpu().numpy().tobytes(), f"e2e result differs in {k}"
            barrier()
            x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            x0.record()
            for _ in range(args.e2e_steps):
                pgt.scan(plan, stat, hcols, minind=5, window_range=(w_lo, w_hi), site_origin=s_lo, out=hout, device=dev)
            x1.record()
            barrier()
            wall = time.perf_counter() - t0
            tt = torch.tensor([x0.elapsed_time(x1), wall * 1e3], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_ms = max(float(tt[0].item()), float(tt[1].item())) / args.e2e_steps
            h2d = bytes_per_site * n_local
            d2h = sum(hout[k].nbytes for k in fields if k in ("sum_a", "sum_b", "fst", "nhet", "nonmissing", "het", "dxy",
                                                               "neffective", "nskip"))
            bb = torch.tensor([float(h2d), float(d2h)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(bb, op=dist.ReduceOp.SUM)
            e2e = {"value": n_total / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(bb[0].item()),
                   "d2h_bytes_per_step": int(bb[1].item()), "ms_per_step": round(e_ms, 3), "steps": args.e2e_steps,
                   "api": "pgt_scan(..., PGT_MEM_HOST): pinned host columns -> 4M-site slabs H2D (double-buffered, "
                          "copy stream) -> level 1 per slab -> level 2 -> D2H of per-window results; window "
                          "positions/labels resolved on the host; per-rank window tables stay on each rank's host",
                   "h2d_gbs": round(float(bb[0].item()) / (e_ms * 1e-3) / 1e9, 2)}
            del hcols, hout
        except Exception as ex:  # report, never fake
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": repr(ex)[:300]}

    # ---- CPU baseline next to it (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu, _ = reference_arm(WORKLOADS["C4"] if fused else wl, min(args.cpu_sample_sites, n_total), 1, 0, with_cli=True)
            cpu["value"] = round(cpu["value"], 1)
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "failed: " + repr(ex)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": int(nw.item()),
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "n_sites": n_total, "contigs": 24,
                       "winsize_sites": W, "stepsize_sites": S, "windows": plan.num_windows, "units": plan.num_units,
                       "unit_sites": unit_sites or 256, "sharding": f"site ranges cut at window starts, halo = W-S = {W - S} sites, {world} shard(s)",
                       "l2": "inputs (>= 6 GB per GPU) exceed the 126 MB L2; no flush needed",
                       "gather": ("NCCL gather of the packed per-window results to rank 0 inside every timed step"
                                  if world > 1 else "single GPU: results stay in HBM")},
            "clocks": clk.summary(),
            "e2e": e2e if e2e is not None else {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": total_launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    a = parse_args()
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
