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


def test_device_mode_scans_can_be_captured_into_a_cuda_graph():
    """With the plan tables resident (pgt_plan_bind_device / pgt_xplan_bind_device, automatic in the Python
    layer) a device-mode scan is kernel launches only: capture once, replay on new column contents."""
    import torch
    import popgenomicstools_b200 as pgt
    dev = torch.device("cuda:0")
    n = 1_000_000
    offs = np.array([0, 600_000, n], np.uint64)
    plan = pgt.WindowPlan(offs, 50000, 10000)
    pos = pgt.synth_pos(3, 0, n, offs, 2, device=dev)
    a, b = pgt.synth_fst(3, 0, n, device=dev)
    score = pgt.synth_score(3, 0, n, device=dev)
    xplan = pgt.ExtremePlan(pos.cpu().numpy(), offs, 100000)
    out = pgt.fst_window(plan, pos, a, b)       # warm-up: binds the tables, sizes the workspaces
    xout = pgt.ihs_window(xplan, pos, score, 2.0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        pgt.fst_window(plan, pos, a, b, out=out)
        pgt.ihs_window(xplan, pos, score, 2.0, out=xout)
    # new contents in the same buffers, replay, compare with a plain call
    a2, b2 = pgt.synth_fst(4, 0, n, device=dev)
    s2 = pgt.synth_score(4, 0, n, device=dev)
    a.copy_(a2); b.copy_(b2); score.copy_(s2)
    for v in list(out.values()) + list(xout.values()):
        v.zero_()
    g.replay()
    torch.cuda.synchronize()
    ref = pgt.fst_window(plan, pos, a2, b2)
    xref = pgt.ihs_window(xplan, pos, s2, 2.0)
    torch.cuda.synchronize()
    for k in ref:
        assert out[k].cpu().numpy().tobytes() == ref[k].cpu().numpy().tobytes(), k
    for k in xref:
        assert xout[k].cpu().numpy().tobytes() == xref[k].cpu().numpy().tobytes(), k
