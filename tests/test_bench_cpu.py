"""bench.py's host-side pieces that need no GPU: the argument contract the driver uses, the workload table read
without importing the package (the reference arm must not load libpgtscan.so), the CPU legs' helpers."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_workloads_load_without_the_package_and_configs_are_named():
    code = ("import sys, bench; w = bench.load_workloads(); "
            "assert 'popgenomicstools_b200' not in sys.modules, 'the reference arm must not import the package'; "
            "print(sorted(w.WORKLOADS)); print([c['name'] for c in bench.CONFIGS])")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "C4" in r.stdout and "C5" in r.stdout
    assert "['C1', 'C2-sparse', 'C2-dense', 'C3-1/1', 'C3-100k', 'C5', 'C5-S1']" in r.stdout


def test_compute_only_port_times_the_oracle_on_preparsed_arrays():
    import bench
    wl = bench.load_workloads()
    _, offs = wl.human_like_contigs(480000, 10000)
    r = bench.compute_only_port(offs, 4, 50000, 10000, 2)
    assert r["kind"] == "port" and r["sites"] == int(offs[-1]) and r["cores"] == 2
    assert r["one_core_sites_per_s"] > 1e6 and r["all_cores_sites_per_s"] >= r["one_core_sites_per_s"] * 0.9


def test_reference_arm_line_on_a_tiny_sample():
    """`bench.py --impl reference` prints ONE JSON line with impl / cpu_baseline / e2e and never touches a GPU."""
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-sites", "2.4e6"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sites/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
