"""CLI start-up floor: where the ~0.3-2 s before the first kernel go.  Every variant runs in a FRESH process
(ctypes on libpgtscan.so, no torch): dlopen, cuInit (pgt_device_count), primary context (pgt_set_device + 1 MB
pgt_device_alloc), first kernel of the library (pgt_synth_fst on 1024 sites + a synchronous copy back: module
load), second kernel (pgt_scan of a tiny plan: the remaining lazily loaded functions).
usage: probe_startup.py            (prints one JSON line per variant)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import ctypes as C, json, os, sys, time
t0 = time.perf_counter()
lib = C.CDLL(os.path.join(sys.argv[1], "popgenomicstools_b200", "libpgtscan.so"))
t1 = time.perf_counter()
n = lib.pgt_device_count()
t2 = time.perf_counter()
lib.pgt_set_device(0)
p = C.c_void_p()
lib.pgt_device_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
lib.pgt_device_alloc(C.byref(p), 1 << 20)
t3 = time.perf_counter()
lib.pgt_synth_fst.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
lib.pgt_memcpy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
buf = (C.c_double * 1024)()
lib.pgt_synth_fst(1, 0, 1024, p, C.c_void_p(p.value + 8192), None)
lib.pgt_memcpy_to_host(buf, p, 8192)
t4 = time.perf_counter()
lib.pgt_synth_fst(1, 0, 1024, p, C.c_void_p(p.value + 8192), None)
lib.pgt_memcpy_to_host(buf, p, 8192)
t5 = time.perf_counter()
print(json.dumps(dict(devices=n, dlopen_ms=round((t1-t0)*1e3,1), cuinit_ms=round((t2-t1)*1e3,1), context_ms=round((t3-t2)*1e3,1),
                      first_kernel_ms=round((t4-t3)*1e3,1), second_kernel_ms=round((t5-t4)*1e3,2), total_ms=round((t5-t0)*1e3,1))))
'''


def run(label, env):
    e = dict(os.environ)
    for k, v in env.items():
        if v is None:
            e.pop(k, None)
        else:
            e[k] = v
    best = None
    for _ in range(3):
        p = subprocess.run([sys.executable, "-c", CHILD, ROOT], capture_output=True, text=True, env=e)
        if p.returncode != 0:
            return dict(label=label, error=p.stderr[-200:])
        r = json.loads(p.stdout.strip().splitlines()[-1])
        if best is None or r["total_ms"] < best["total_ms"]:
            best = r
    best["label"] = label
    return best


if __name__ == "__main__":
    for label, env in (("one visible device (what the tools do), lazy module loading (CUDA 12 default)", {"CUDA_VISIBLE_DEVICES": "0"}),
                       ("one visible device, CUDA_MODULE_LOADING=EAGER", {"CUDA_VISIBLE_DEVICES": "0", "CUDA_MODULE_LOADING": "EAGER"}),
                       ("all devices visible, lazy", {"CUDA_VISIBLE_DEVICES": None}),
                       ("one visible device, CUDA_DEVICE_MAX_CONNECTIONS=1", {"CUDA_VISIBLE_DEVICES": "0", "CUDA_DEVICE_MAX_CONNECTIONS": "1"})):
        print(json.dumps(run(label, env)), flush=True)
