#!/bin/bash
# Host-side planning code of the hits scan (block plan, filter tables, direct-path enumeration, lock-free hash
# fill: csrc/scan_hits.cu, csrc/scan.cu) under ThreadSanitizer and AddressSanitizer + UBSan.  No GPU needed: the
# drivers call the host-only test hooks of include/tfbs_cuda.h.  (compute-sanitizer, which would cover the
# DEVICE code, is closed on the GPU pool.)  Usage: bash tests/sanitizers/run.sh   -> two binaries in /tmp.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$HERE/../.."
SRC="$ROOT/discover_tfbs_b200/csrc"
OUT="${TMPDIR:-/tmp}/tfbs_sanitizers"
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
COMMON="-gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 --expt-relaxed-constexpr -ccbin /usr/bin/g++ -I$ROOT/include"
FILES="$SRC/scan_hits.cu $SRC/scan.cu $SRC/api.cu $SRC/encode.cu"
$NVCC $COMMON -Xcompiler -fsanitize=thread,-fPIC,-fno-omit-frame-pointer -o "$OUT/plan_tsan" "$HERE/plan_threads.cpp" $FILES -Xlinker -ltsan
$NVCC $COMMON -Xcompiler -fsanitize=address,-fsanitize=undefined,-fPIC,-fno-omit-frame-pointer -o "$OUT/plan_asan" "$HERE/plan_cases.cpp" $FILES -Xlinker -lasan -Xlinker -lubsan
echo "== ThreadSanitizer: parallel planning (direct-path count pass, lock-free hash fill, lane-range block forms)"
TSAN_OPTIONS="halt_on_error=1" "$OUT/plan_tsan"
echo "== AddressSanitizer + UBSan: 14 libraries (1 - 800 motifs, lengths 1 - 32, -inf entries, +-inf / NaN thresholds), every filter form"
ASAN_OPTIONS="protect_shadow_gap=0:detect_leaks=0" UBSAN_OPTIONS="halt_on_error=1" "$OUT/plan_asan"
echo "sanitizers: clean"
