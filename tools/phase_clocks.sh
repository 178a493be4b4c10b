#!/bin/bash
# usage (on the GPU box): tools/phase_clocks.sh <config> <N...>   -> per-phase cycles of warp 0 (development build)
cd "$(dirname "$0")/.."
cp libzombsole_b200/csrc/libzs_b200.so /tmp/libzs_b200.keep
(cd libzombsole_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
  -Xcompiler -fvisibility=hidden --fmad=false -cudart static -DZS_PHASE_CLOCKS -o libzs_b200.so zs_b200.cu 2>/dev/null)
cfg=$1; shift
for l in 32 16; do echo "lanes $l NO_TMA=$ZS_NO_TMA"; ZS_LANES_PER_ENV=$l timeout 120 python tools/probe_lanes.py $cfg "$@" 2>&1 | tail -n 12; done
cp /tmp/libzs_b200.keep libzombsole_b200/csrc/libzs_b200.so
