#!/bin/bash
# usage (GPU box): tools/checked_build.sh [pytest args]  -> the GPU parity tests and the smoke runs on a -DZS_CHECKS build:
# every shared-memory view access (grid, lists, slot arrays, candidate lists) is bounds-checked on the device and a
# violation traps.  (compute-sanitizer is closed on this pool; this is the substitute for its memcheck pass.)
cd "$(dirname "$0")/.."
cp libzombsole_b200/csrc/libzs_b200.so /tmp/libzs_b200.keep
(cd libzombsole_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
  -Xcompiler -fvisibility=hidden --fmad=false -cudart static -Xcompiler -fopenmp -lgomp -DZS_CHECKS -o libzs_b200.so zs_b200.cu 2>&1 | grep -E "error" | head -5)
timeout 1500 python tools/sanitize_smoke.py 2>&1 | tail -8
timeout 2400 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -6
cp /tmp/libzs_b200.keep libzombsole_b200/csrc/libzs_b200.so
