#!/bin/bash
# The host build of the code shared by host and device (csrc/chess_core.cuh, movegen_warp.cuh with its lanes simulated on the host,
# ssl_core.cuh) under AddressSanitizer + UndefinedBehaviorSanitizer, driven by the same tests as the plain host check:
#     bash tests/hostcheck/run_sanitized.sh [log file]
# compute-sanitizer is closed on the GPU pool (profiles/r02_compute_sanitizer_closed_on_pool.log); this covers out-of-bounds accesses,
# shifts past the word size, signed overflow and misaligned accesses in that shared code, not the kernels' own indexing.
set -u
cd "$(dirname "$0")/../.."
LOG=${1:-/dev/stdout}
ASAN=$(gcc -print-file-name=libasan.so)
M0_HOSTCHECK_SANITIZE=1 LD_PRELOAD="$ASAN" ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
  python -m pytest tests/test_hostcheck.py -q -x -p no:cacheprovider > "$LOG" 2>&1
rc=$?
echo "sanitized hostcheck rc=$rc" >> "$LOG"
exit $rc
