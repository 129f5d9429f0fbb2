#!/bin/bash
# The library GPU tests against the -DPGT_BOUNDS build: every index a kernel forms is checked against its extent and a
# violation traps.  (The CLI tests are skipped: the binaries link the ordinary library.)
set -e
cd "$(dirname "$0")/.."
make -s -C popgenomicstools_b200/csrc bounds
PGT_LIB=$PWD/popgenomicstools_b200/libpgtscan_bounds.so python -m pytest tests -m gpu -x -q \
    --deselect tests/test_fullscale_gpu.py \
    --ignore tests/test_cli_gpu.py --ignore tests/test_extreme_cli_gpu.py --ignore tests/test_binding_gpu.py "$@"
