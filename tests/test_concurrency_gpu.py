"""The library is re-entrant and stream-ordered (SURVEY.md section 8b "Threading"): scans issued from
several host threads, each on its own CUDA stream with its own plan / workspace, give the same bits
as the same scans issued one after the other."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_scans_from_threads_match_serial():
    import torch
    import popgenomicstools_b200 as pgt
    dev = torch.device("cuda:0")
    jobs = []
    for i, (n, W, S) in enumerate([(3_000_000, 5000, 1000), (2_000_000, 1000, 100), (1_500_000, 64, 1), (2_500_000, 777, 13)]):
        offs = np.array([0, n // 3, n], np.uint64)
        a, b = pgt.synth_fst(10 + i, 0, n, device=dev)
        g = pgt.synth_het(10 + i, 0, n, device=dev)
        pos = pgt.synth_pos(10 + i, 0, n, offs, 3, device=dev)
        score = pgt.synth_score(10 + i, 0, n, device=dev)
        xplan = pgt.ExtremePlan(pos.cpu().numpy(), offs, 20000 + 1000 * i)
        jobs.append(dict(plan=pgt.WindowPlan(offs, W, S), hplan=pgt.WindowPlan(offs, W, S, unit_sites=4096), xplan=xplan,
                         pos=pos, a=a, b=b, g=g, score=score))
    torch.cuda.synchronize()

    def run(j, out):
        with torch.cuda.stream(torch.cuda.Stream(dev)):
            for _ in range(10):
                r1 = pgt.fst_window(j["plan"], j["pos"], j["a"], j["b"])
                r2 = pgt.het_window(j["hplan"], j["pos"], j["g"])
                r3 = pgt.ihs_window(j["xplan"], j["pos"], j["score"], 2.0)
            torch.cuda.current_stream().synchronize()
            out.append({**{"fst_" + k: v.cpu().numpy() for k, v in r1.items()}, **{"het_" + k: v.cpu().numpy() for k, v in r2.items()},
                        **{"ihs_" + k: v.cpu().numpy() for k, v in r3.items()}})

    serial = []
    for j in jobs:
        o = []
        run(j, o)
        serial.append(o[0])
    outs = [[] for _ in jobs]
    errs = []

    def guarded(j, o):
        try:
            run(j, o)
        except Exception as e:  # surfaced below: an exception in a thread must fail the test
            errs.append(e)

    th = [threading.Thread(target=guarded, args=(j, o)) for j, o in zip(jobs, outs)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for s, o in zip(serial, outs):
        assert len(o) == 1
        for k in s:
            assert o[0][k].tobytes() == s[k].tobytes(), k
