"""The reference-side binding (tests/integration/sitewindow_binding.cpp) on the
GPU: the reference's own main + argument parser + stringstream line loop, with calcWindow and its flush
triggers replaced by pgt_plan_create + pgt_scan through the C ABI, prints the committed transcripts of
the unmodified binaries."""
import os

import pytest

import cli_util as U
import parity as P
from test_cli_gpu import write_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_binding_prints_the_reference_transcripts(golden_cases, tmp_path):
    ties = ran = 0
    for i, c in enumerate(golden_cases):
        if c["tool"] not in ("fstWindow", "hetWindow") or i % 12:
            continue  # each process pays 1-3 s of CUDA start-up
        exe = os.path.join(U.ROOT, "oracle", "_ref", c["tool"] + "_pgt")
        if not os.access(exe, os.X_OK):
            pytest.skip("oracle/_ref/*_pgt not built (make -C oracle binding needs the reference sources)")
        d = tmp_path / f"c{i}"
        d.mkdir()
        write_case(c, str(d))
        rc, out, err = U.run(exe, c["argv"], cwd=str(d))
        assert (rc, err) == (c["rc"], c["stderr"]), (i, c["argv"], err)
        ties += P.rows_match_modulo_ties(out.splitlines(), c["stdout"].splitlines(), {4})
        ran += 1
    assert ran >= 6 and ties <= 1
