#!/bin/bash
# ThreadSanitizer run of the emulated kernels (see tools/tsan_emulated.py).  Output: the TSan report on stderr, a one-line summary last.
set -e
cd "$(dirname "$0")/.."
out=${TMPDIR:-/tmp}/libhostcheck_tsan.so
g++ -O1 -g -std=c++20 -fPIC -ffp-contract=off -march=x86-64-v3 -fsanitize=thread -shared -I${CUDA_HOME:-/usr/local/cuda}/include \
    -o $out tests/hostcheck/hostcheck.cpp tests/hostcheck/hostcheck_warp.cpp prt_b200/csrc/bvh_build.cpp -lpthread
PRT_HOSTCHECK_TSAN=$out LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0" \
    python tools/tsan_emulated.py
