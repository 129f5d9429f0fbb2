#!/bin/bash
# Host-side hygiene run (no GPU): builds libpgtscan.so and the five CLIs with
# -fsanitize=address,undefined into build/asan/ and runs the CPU tests that exercise the host code
# (closed-form planner, extreme-scan window bookkeeping, ABI, text parsers / .pgtc packing, argv handling).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/build/asan"
rm -rf "$OUT"; mkdir -p "$OUT/pkg"
cd "$ROOT/popgenomicstools_b200/csrc"
SAN="-fsanitize=address,-fsanitize=undefined,-fno-omit-frame-pointer"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC,$SAN -cudart static -shared \
    pgt_scan.cu pgt_sharded.cu pgt_upload.cu pgt_extreme.cu pgt_synth.cu pgt_plan.cpp -o "$OUT/libpgtscan.so"
cp -r "$ROOT/popgenomicstools_b200" "$ROOT/tests" "$ROOT/oracle" "$ROOT/include" "$OUT/pkg/"
cp "$OUT/libpgtscan.so" "$OUT/pkg/popgenomicstools_b200/libpgtscan.so"
mkdir -p "$OUT/bin"
for t in "fstWindow -DPGT_TOOL_FST tools/sitewindow_main.cpp" "hetWindow -DPGT_TOOL_HET tools/sitewindow_main.cpp" \
         "dxyWindow -DPGT_TOOL_DXY tools/dxywindow_main.cpp" "ihsWindow -DPGT_TOOL_IHS tools/extremewindow_main.cpp" \
         "xpehhWindow -DPGT_TOOL_XPEHH tools/extremewindow_main.cpp"; do
    set -- $t
    g++ -O1 -g -std=c++17 -Wall -pthread -fsanitize=address,undefined -fno-omit-frame-pointer $2 $3 -o "$OUT/bin/$1" \
        -L"$OUT" -lpgtscan -Wl,-rpath,"$OUT" -lz -ldl -lrt
done
cd "$OUT/pkg"
export ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
LD_PRELOAD="$(gcc -print-file-name=libasan.so)" python -m pytest tests/test_plan_cpu.py tests/test_extreme_plan_cpu.py tests/test_abi_cpu.py -x -q
cd "$ROOT"  # the CLI tests import the normal package; only the CLI binaries are the sanitizer builds
PGT_TEST_BIN="$OUT/bin" python -m pytest tests/test_colfile_cpu.py tests/test_extreme_cli_cpu.py tests/test_cli_cpu.py tests/test_parser_fuzz_cpu.py -x -q
echo "asan/ubsan host run: clean"
