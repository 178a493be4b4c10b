#!/bin/bash
# usage (GPU box): tools/trace_launch.sh [K ...]  -> where the time of a K-step launch goes, from per-warp global-timer
# stamps (development build, -DZS_TRACE; TRACE_PY=<script> runs another reader, e.g. tools/trace_host_step.py): kernel entry spread, prologue, each of the first steps, exit spread.
cd "$(dirname "$0")/.."
cp libzombsole_b200/csrc/libzs_b200.so /tmp/libzs_b200.keep
(cd libzombsole_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
  -Xcompiler -fvisibility=hidden -Xcompiler -fopenmp -lgomp --fmad=false -cudart static -DZS_TRACE -o libzs_b200.so zs_b200.cu 2>/dev/null)
timeout 300 python ${TRACE_PY:-tools/trace_launch.py} "$@"
cp /tmp/libzs_b200.keep libzombsole_b200/csrc/libzs_b200.so
